/* The C ABI consumed from plain C (C99), no Python and no C++: include/vmasr_b200.h must be a valid C header and the
 * shared library must link and answer without a GPU.  Built and run by tests/test_abi_cpu.py::test_header_is_valid_c. */
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "vmasr_b200.h"

int main(void) {
    if (vmasr_abi_version() != VMASR_ABI_VERSION) return 1;
    /* workspace sizing: one chunk needs none, 128 chunks need at least 16 bytes per (batch, channel, chunk) */
    if (vmasr_scan_workspace_bytes(4, 8, 2048, 1) != 0) return 2;
    if (vmasr_scan_workspace_bytes(4, 8, 262144, 1) < (uint64_t)4 * 8 * 128 * 16) return 3;

    /* plan the full-resolution layer of the 48 kHz config: B = 4, D = 8, L = 262144, 4 groups (fake, aligned pointers) */
    static char arena[1 << 16];
    vmasr_scan_params p;
    memset(&p, 0, sizeof p);
    p.u = p.delta = p.B = p.C = p.out = arena;
    p.A = (const float *)arena;
    p.x = (float *)arena;
    p.workspace = arena;
    p.workspace_bytes = vmasr_scan_workspace_bytes(4, 8, 262144, 1);
    p.batch = 4; p.dim = 8; p.seqlen = 262144; p.dstate = 1; p.ngroups = 4;
    p.u_batch_stride = p.delta_batch_stride = p.out_batch_stride = 8LL * 262144;
    p.u_d_stride = p.delta_d_stride = p.out_d_stride = 262144;
    p.A_d_stride = 1; p.A_dstate_stride = 1;
    p.B_batch_stride = p.C_batch_stride = 4LL * 262144;
    p.B_group_stride = p.C_group_stride = 262144;
    p.B_dstate_stride = p.C_dstate_stride = 262144;
    p.io_dtype = VMASR_F32; p.delta_softplus = 1; p.device = 0;
    int32_t plan[6];
    if (vmasr_scan_plan(&p, 0, plan) != 0) { fprintf(stderr, "%s\n", vmasr_last_error()); return 4; }
    printf("grid %d family %d channels_per_tile %d chunks %d\n", plan[0], plan[1], plan[2], plan[4]);
    if (plan[4] != 128 || plan[1] != 2 || plan[0] != 4 * 4 * 128 * plan[3]) return 5;

    /* an argument error comes back as a code and a message, never as an exception or an abort */
    p.ngroups = 3;
    if (vmasr_scan_plan(&p, 0, plan) == 0) return 6;
    if (strstr(vmasr_last_error(), "n_groups") == NULL) return 7;
    return 0;
}
