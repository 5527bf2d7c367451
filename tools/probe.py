"""Hardware probes run on the B200 box: FP32 FMA-pipe peak (scalar FFMA vs packed FFMA2) and the
standalone pairwise neighbour-alignment pass.  Prints one JSON object."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evacuation_b200 import _native as nat


def main():
    lib = nat.load()
    out = {}
    ms, fl = C.c_float(), C.c_double()
    for packed in (0, 1):
        best = 0.0
        for _ in range(3):
            nat.check(lib.evac_probe_fma(0, packed, 20000, C.byref(ms), C.byref(fl)))
            best = max(best, fl.value / (ms.value * 1e-3) / 1e12)
        out["fma_tflops_packed" if packed else "fma_tflops_scalar"] = best
    for mode, name in ((2, "ffma2_distinct_operands"), (3, "fadd2_broadcast_operand"), (4, "fmul2"), (5, "pairwise_mix_registers_only"),
                       (6, "ffma2_acc_a64_bcast32"), (7, "fsetp_predicated_fadd2"), (8, "ffma2_acc_a_a"), (9, "ffma2_acc_shared_a"),
                       (10, "ffma2_acc_three_distinct"), (11, "ffma_scalar_three_distinct")):
        best = 0.0
        for _ in range(3):
            nat.check(lib.evac_probe_fma(0, mode, 2000 if mode == 5 else 20000, C.byref(ms), C.byref(fl)))
            best = max(best, fl.value / (ms.value * 1e-3) / 1e12)
        out["fma_pipe_tflops_equiv_" + name] = best
    pairs = C.c_double()
    for (E, n, reps) in ((4096, 60, 200), (65536, 60, 50), (148 * 28 * 4, 64, 200), (148 * 32 * 4, 64, 200), (148 * 32 * 4, 60, 200)):
        best = 0.0
        for _ in range(3):
            nat.check(lib.evac_probe_pairwise(0, E, n, reps, C.byref(ms), C.byref(pairs)))
            best = max(best, pairs.value / (ms.value * 1e-3))
        out[f"pairs_per_s_E{E}_N{n}"] = best
        out[f"pair_tflops_E{E}_N{n}"] = best * 8 / 1e12
    print(json.dumps(out))


if __name__ == "__main__":
    main()
