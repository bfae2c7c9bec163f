"""Builds tuning variants of libc2c_b200.so under build/variants/<name>.so (macros of csrc/c2c_contract.cuh)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cell2cell_b200._build import ROOT, build  # noqa: E402

VARIANTS = {
    "u8_b2": ["-DC2C_PC_UNROLL=8", "-DC2C_PC_MINB=2", "-DC2C_FC_UNROLL=8"],
    "u4_b2": ["-DC2C_PC_UNROLL=4", "-DC2C_PC_MINB=2", "-DC2C_FC_UNROLL=4", "-DC2C_FC_MINB=2", "-DC2C_FC_THREADS_LO=256"],
    "u8_b3": ["-DC2C_PC_UNROLL=8", "-DC2C_PC_MINB=3", "-DC2C_FC_UNROLL=8", "-DC2C_FC_MINB=2", "-DC2C_FC_THREADS_LO=256"],
}
if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "build", "variants"), exist_ok=True)
    names = sys.argv[1:] or list(VARIANTS)
    for n in names:
        print(build(extra_flags=VARIANTS[n], out_path=os.path.join(ROOT, "build", "variants", n + ".so"), verbose=True))
