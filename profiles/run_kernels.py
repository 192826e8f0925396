"""Profiling driver: K1 (residual + Jacobian) and K2 (normal-equation accumulation) alone on a synthetic
flight, for `ncu -k regex:<kernel> -c 1 --set full ... python profiles/run_kernels.py` (B200_PROFILING.md).
Prints the CUDA-event times (never quote a time measured under ncu)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cams', type=int, default=64)
    ap.add_argument('--det', type=int, default=1000000)
    ap.add_argument('--coef', type=int, default=200000)
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('--solve-steps', type=int, default=0, help='also run this many LM steps (linear-solve kernels)')
    a = ap.parse_args()
    import bench
    from mvus_b200 import _cabi
    from mvus_b200.problem import FlatProblem
    fl = bench.make_workload(a.cams, a.det, a.coef)
    fp = FlatProblem(fl, fl.numCam, **bench.BA_KW)
    hd = _cabi.Handle(fp, ftol=0.0, xtol=0.0, gtol=0.0, max_nfev=max(a.solve_steps, 1) + 1)
    k1 = hd.time_resjac(fp.x0, reps=a.reps)
    k2 = hd.time_accumulate(reps=a.reps)
    print('N %d  K1 %.3f ms  K2 %.3f ms' % (fp.N, k1, k2))
    if a.solve_steps:
        x, r, st = hd.solve(fp.x0, want_r=False)
        print(st.as_dict())
    hd.close()


if __name__ == '__main__':
    main()
