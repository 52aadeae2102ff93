"""Runs tests/csrc/mma_probe.cu (compiled on the spot with nvcc) and prints cycles per tcgen05.mma for a set of operand
configurations -- evidence for DESIGN.md 4.3 (what bounds the MMA phase of k_net_tc).  Bring-up tool, not a test."""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = "/tmp/libmmaprobe.so"
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-shared", "-Xcompiler", "-fPIC",
                       "-I", os.path.join(ROOT, "lightzero_b200", "csrc"), os.path.join(ROOT, "tests", "csrc", "mma_probe.cu"), "-o", so])
lib = ctypes.CDLL(so)


class Cfg(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("N", "n_mma", "n_acc", "switch_every", "a_shift_rows", "a_lbo16", "swz_a", "swz_b", "n_ksteps", "mix_n2", "mix_group", "mix_d2_off", "warp_issue")]


NM = 384
cases = []
for wi, tag in ((0, "lane0 "), (1, "elect ")):
    cases += [
        (tag + "N=64  one accumulator, aligned A", Cfg(64, NM, 1, NM, 0, 128, 0, 0, 4, 0, 1, 0, wi)),
        (tag + "N=64  A shifted by 7 rows, plane stride 400 rows (net_tc)", Cfg(64, NM, 1, NM, 7, 400, 0, 0, 4, 0, 1, 0, wi)),
        (tag + "N=64  3 accumulators, switch every MMA", Cfg(64, NM, 3, 1, 0, 128, 0, 0, 4, 0, 1, 0, wi)),
        (tag + "N=64  3 accumulators, switch every 12 (net_tc order)", Cfg(64, NM, 3, 12, 0, 128, 0, 0, 4, 0, 1, 0, wi)),
        (tag + "N=32  one accumulator", Cfg(32, NM, 1, NM, 0, 128, 0, 0, 4, 0, 1, 0, wi)),
        (tag + "N=128 one accumulator", Cfg(128, NM, 1, NM, 0, 128, 0, 0, 4, 0, 1, 0, wi)),
        (tag + "N=256 one accumulator", Cfg(256, NM, 1, NM, 0, 128, 0, 0, 2, 0, 1, 0, wi)),
        (tag + "N=64  A SWIZZLE_128B, shifted 7 rows", Cfg(64, NM, 1, NM, 7, 128, 1, 0, 4, 0, 1, 0, wi)),
        (tag + "N=64  A and B SWIZZLE_128B", Cfg(64, NM, 1, NM, 0, 128, 1, 1, 4, 0, 1, 0, wi)),
    ]
cases += [
    ("elect N=128/N=64 alternating groups of 4, SAME accumulator columns (the fold as first written)", Cfg(128, NM, 1, NM, 0, 400, 0, 0, 4, 64, 4, 0, 1)),
    ("elect N=128/N=64 alternating groups of 12, same columns", Cfg(128, NM, 1, NM, 0, 400, 0, 0, 4, 64, 12, 0, 1)),
    ("elect N=128/N=64 alternating groups of 4, N=64 into OTHER columns (+256)", Cfg(128, NM, 1, NM, 0, 400, 0, 0, 4, 64, 4, 256, 1)),
    ("elect N=128/N=64 alternating groups of 4, N=64 into the upper half (+64)", Cfg(128, NM, 1, NM, 0, 400, 0, 0, 4, 64, 4, 64, 1)),
    ("elect N=128/N=128 alternating groups of 4 (two N=128 passes), same columns", Cfg(128, NM, 1, NM, 0, 400, 0, 0, 4, 128, 4, 0, 1)),
    ("elect N=64/N=64 alternating groups of 4, same columns (control)", Cfg(64, NM, 1, NM, 0, 400, 0, 0, 4, 64, 4, 0, 1)),
    ("elect N=64/N=64 alternating groups of 4, other columns +64 (control)", Cfg(64, NM, 1, NM, 0, 400, 0, 0, 4, 64, 4, 64, 1)),
    ("elect N=128 3 accumulators switch every 4", Cfg(128, NM, 3, 4, 0, 400, 0, 0, 4, 0, 1, 0, 1)),
    ("elect N=192 one accumulator", Cfg(192, NM, 1, NM, 0, 400, 0, 0, 4, 0, 1, 0, 1)),
    ("elect N=96 one accumulator", Cfg(96, NM, 1, NM, 0, 400, 0, 0, 4, 0, 1, 0, 1)),
]
arr = (Cfg * len(cases))(*[c for _, c in cases])
out = (ctypes.c_ulonglong * len(cases))()
rc = lib.mma_probe_run(arr, len(cases), out)
print("configurations measured:", rc, "of", len(cases))
print(f"{'configuration':66s} cycles/MMA   math floor (128*N/256)")
for (name, c), cyc in zip(cases, out):
    floor = 128 * c.N / 256 if not c.mix_n2 else 128 * (c.N + c.mix_n2) / 2 / 256
    print(f"{name:100s} {cyc / c.n_mma:8.1f}     {floor:6.0f}")
