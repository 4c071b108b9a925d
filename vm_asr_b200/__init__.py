"""Import alias: the package lives in ``vm-asr_b200/`` (not a Python identifier), so this stub package
extends its search path to that directory.  ``import vm_asr_b200.scan`` loads ``vm-asr_b200/scan.py``."""
import os as _os

__path__.append(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "vm-asr_b200"))

from ._lib import library_path, load_library  # noqa: E402,F401
