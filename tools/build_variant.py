"""Builds the library with extra nvcc flags into variants/NAME.so (not loaded by the package;
swap it in on a scratch copy to compare):  python tools/build_variant.py bounds -DLBL_DEBUG_BOUNDS

Knobs read by the kernels: LBL_DEBUG_BOUNDS (device-side index asserts), LBL_STAGE_LINES,
LBL_STAGES (staging ring of the far-field kernel), LBL_CELL_WARPS, LBL_CELL_RESIDENT,
LBL_NEAR_RESIDENT (block shapes / launch bounds)."""
import subprocess
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pylbl_b200 import build as b

name, flags = sys.argv[1], sys.argv[2:]
out = Path(__file__).resolve().parent.parent / "variants"
out.mkdir(exist_ok=True)
cmd = [b._nvcc()] + b.NVCC_FLAGS + flags + ["-o", str(out / f"{name}.so")] + [str(b.CSRC / s) for s in b.SOURCES] + ["-ldl"]
proc = subprocess.run(cmd, capture_output=True, text=True)
if proc.returncode:
    sys.stderr.write(proc.stderr[-4000:])
    sys.exit(proc.returncode)
print(out / f"{name}.so")
