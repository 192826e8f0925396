#!/usr/bin/env python
"""Benchmark of the mvus BA hot path (contract: see the task statement / DESIGN.md section 7).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--cams C --det D --coef NCOEF] [--no-cpu-baseline]

A step = one Levenberg-Marquardt iteration of the bundle adjustment (linear solve(s) for the
damped step, trial residual evaluation, and -- when the step is accepted -- residual+Jacobian
and normal-equation accumulation at the new point) on the synthetic flight named in
config.workload.  value = detections x LM iterations per second over all GPUs, inputs resident
in HBM, device-timed with CUDA events inside the library; e2e = the same through Scene.BA with
host buffers (H2D of the detections and D2H of the result inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'BA LM throughput (detections x LM iterations / s; resid+Jacobian Mdet/s and LM iters/s alongside)'
UNIT = 'Mdet/s'
BA_KW = dict(rs=True, motion_reg=True, motion_weights=1e4)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=8)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--cams', type=int, default=64)
    ap.add_argument('--det', type=int, default=1000000, help='detections per camera')
    ap.add_argument('--coef', type=int, default=200000, help='spline coefficients per axis')
    ap.add_argument('--sample-cams', type=int, default=7)
    ap.add_argument('--sample-det', type=int, default=3000)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--workload', default='cfg4', choices=['cfg4', 'cfg2', 'cfg3', 'cfg5'],
                    help='BASELINE.json config: cfg4 (default, the one the metric is quoted on), cfg2 7x100k RS+F, '
                         'cfg3 = cfg2 + opt_calib + KE, cfg5 = batch of independent 7-camera problems')
    ap.add_argument('--problems', type=int, default=1024, help='cfg5: number of independent problems')
    ap.add_argument('--batch-threads', type=int, default=8, help='cfg5: problems in flight per GPU (host threads)')
    return ap.parse_args()


def make_workload(cams, det, coef, opt_calib=False, motion_type='F', frames_per_knot=None):
    from concurrent.futures import ThreadPoolExecutor
    from mvus_b200 import synth
    # simulate cameras in parallel threads (NumPy releases the GIL in the heavy ufuncs)
    orig = synth.simulate_detections
    pool = ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1))
    futures = []

    def deferred(*a, **k):
        # each camera needs its own generator state to be reproducible under threading
        a = list(a)
        a[6] = np.random.default_rng(1000 + len(futures))
        fut = pool.submit(orig, *a, **k)
        futures.append(fut)
        return fut
    synth.simulate_detections = deferred
    try:
        fl, truth = synth.make_flight(nc=cams, det_per_cam=det, n_coef=None if frames_per_knot else coef,
                                      frames_per_knot=frames_per_knot or 15.0, rolling_shutter=True,
                                      distortion=True, opt_calib=opt_calib, motion_type=motion_type,
                                      motion_weights=1e4, uncovered=0.0)
    finally:
        synth.simulate_detections = orig
    fl.detections = [f.result() for f in fl.detections]
    pool.shutdown()
    return fl


def workload_name(a):
    if a.workload == 'cfg2':
        return 'cfg2: synthetic 7-camera x 100000-detection flight, 15 frames/knot, rolling shutter + motion F (w=1e4)'
    if a.workload == 'cfg3':
        return 'cfg3: cfg2 with opt_calib (15 camera unknowns) and motion KE (w=1e2)'
    return ('synthetic %d-camera x %d-detection flight, %d spline coefficients/axis, rolling shutter + '
            'motion F (w=1e4), fixed calibration' % (a.cams, a.det, a.coef))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for k, n in enumerate(names) if any(len(r) > 3 + k and r[3 + k] == 'Active' for r in self.rows)]
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


def sample_flight(a):
    """The bounded CPU sample of the workload: same settings (rolling shutter, distortion, motion F,
    w = 1e4), a.sample_cams cameras x a.sample_det detections -- the size at which the UNMODIFIED
    reference still runs (its jac_BA pattern is a dense int64 m x n array, common.py:610)."""
    from mvus_b200 import synth
    fl, _ = synth.make_flight(nc=a.sample_cams, det_per_cam=a.sample_det, rolling_shutter=True,
                              distortion=True, motion_type='F', motion_weights=1e4, uncovered=0.0)
    return fl


def cpu_reference_run(a, steps, warmup):
    """The reference's CPU path on a bounded sample of the workload, all host threads BLAS wants.
    kind "reference": the UNMODIFIED reference (oracle/_ref, the verbatim copy oracle/make_ref.py makes)
    -- Scene.BA = jac_BA pattern + scipy least_squares(TRF, LSMR, 2-point FD) exactly as common.py:670;
    kind "port": the oracle's restatement of that call, only when oracle/_ref did not travel.
    BASELINE.md section 3 legs: (1) error_BA residual Mdet/s, (2) approx_derivative residual+Jacobian
    Mdet/s, (3) LM iterations/s.  Returns (Mdet/s of leg 3, info dict, seconds, iterations)."""
    import contextlib
    import io
    from oracle import ba_oracle, ref_shim
    from scipy.optimize._numdiff import approx_derivative, group_columns
    fl = sample_flight(a)
    N = int(sum(d.shape[1] for d in fl.detections))
    kw = dict(BA_KW)
    quiet = contextlib.redirect_stdout(io.StringIO())
    if ref_shim.available():
        kind = 'reference'
        t0 = time.perf_counter()
        with quiet:
            fn, x0, A, _ = ref_shim.capture_ba(ref_shim.to_reference_scene(fl), fl.numCam, **kw)
        t_pat = time.perf_counter() - t0                  # jac_BA (compute_visibility + pattern), common.py:665

        def solve(max_nfev):
            ref = ref_shim.to_reference_scene(fl)
            with quiet:
                return ref.BA(fl.numCam, max_iter=max_nfev, **kw)
    else:
        kind = 'port'
        prob = ba_oracle.Problem(fl, fl.numCam, **kw)
        x0, fn = prob.x0, prob.residual
        t0 = time.perf_counter()
        A = prob.pattern_near3(prob.x0)
        t_pat = time.perf_counter() - t0

        def solve(max_nfev):
            return prob.shipped_solve(prob.x0, max_nfev=max_nfev, pattern=A)
    n, m = len(x0), len(fn(x0))
    ts = []
    for _ in range(5):                                    # leg 1
        t0 = time.perf_counter()
        f0 = fn(x0)
        ts.append(time.perf_counter() - t0)
    t_res = float(np.median(ts))
    groups = group_columns(A)
    t0 = time.perf_counter()                              # leg 2: what least_squares does per Jacobian
    approx_derivative(fn, x0, method='2-point', f0=f0, sparsity=(A, groups))
    t_jac = time.perf_counter() - t0 + t_res
    if warmup:
        solve(2)
    t0 = time.perf_counter()                              # leg 3
    res = solve(steps + 1)
    dt = time.perf_counter() - t0 - (t_pat if kind == 'reference' else 0.0)     # (Scene.BA rebuilds the pattern)
    iters = max(res.nfev - 1, 1)
    val = N * iters / dt / 1e6
    info = {'value': val, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': kind,
            'sample': '%d cameras x %d detections (N=%d, n=%d, m=%d), same settings as the workload; %d LM iterations '
                      '(nfev %d, njev %d, status %d) in %.2f s; jac_BA pattern %.2f s not included' % (
                          a.sample_cams, a.sample_det, N, n, m, iters, res.nfev, res.njev, res.status, dt, t_pat),
            'lm_iters_per_s': iters / dt, 'final_cost': float(res.cost),
            'resid_mdet_per_s': N / t_res / 1e6, 'resid_jac_mdet_per_s': N / t_jac / 1e6,
            'fd_colour_groups': int(groups.max()) + 1, 'pattern_s': t_pat}
    return val, info, dt, iters


def gpu_on_sample(a, steps):
    """Same-config pair of the CPU arm: the sample flight through the public API on this GPU."""
    from mvus_b200 import ba
    fl = sample_flight(a)
    N = int(sum(d.shape[1] for d in fl.detections))
    import contextlib
    import io
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        ba.bundle_adjust(sample_flight(a), a.sample_cams, max_iter=3, ftol=0.0, xtol=0.0, gtol=0.0, **BA_KW)   # warm-up
        t0 = time.perf_counter()
        res = ba.bundle_adjust(fl, fl.numCam, max_iter=steps + 1, ftol=0.0, xtol=0.0, gtol=0.0, **BA_KW)
        dt = time.perf_counter() - t0
    it = max(res.nfev - 1, 1)
    out = {'workload': '%d cameras x %d detections' % (a.sample_cams, a.sample_det), 'steps': it,
           'gpu_e2e_value': N * it / dt / 1e6, 'gpu_e2e_seconds': dt,
           'gpu_value': N * it / (res.stats['ms_total'] / 1e3) / 1e6, 'gpu_final_cost': float(res.cost), 'unit': UNIT}
    return out


def run_cfg5(a, rank, world, local):
    """Batch of independent problems (BASELINE config 5), partitioned across ranks."""
    import torch
    import torch.distributed as dist
    from mvus_b200 import ba, batch, synth
    torch.cuda.set_device(local)
    ba.DEVICE = local
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    mine = batch.my_problems(a.problems, rank, world)
    scenes = {p: synth.make_flight(nc=7, det_per_cam=5000, seed=7 * p, rolling_shutter=True, distortion=True,
                                   motion_type='F', motion_weights=1e4, uncovered=0.0)[0] for p in mine}
    ndet = sum(sum(d.shape[1] for d in s.detections) for s in scenes.values())
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        for p in mine[:2]:                                            # warm-up on copies of two problems
            synth.make_flight(nc=7, det_per_cam=5000, seed=7 * p, rolling_shutter=True, distortion=True,
                              motion_type='F', motion_weights=1e4, uncovered=0.0)[0].BA(7, max_iter=a.steps + 1, **BA_KW)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = batch.solve_scenes(scenes, numCam=7, threads=a.batch_threads, max_iter=a.steps + 1, **BA_KW)
        torch.cuda.synchronize()
    tot = torch.tensor([time.perf_counter() - t0, float(ndet), float(sum(r.nfev - 1 for r in res.values())),
                        sum(r.stats['ms_total'] for r in res.values()), float(sum(r.stats['launches'] for r in res.values()))],
                       dtype=torch.float64, device='cuda')
    mx = tot.clone()
    if world > 1:
        dist.all_reduce(tot)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    tot, mx = tot.cpu().tolist(), mx.cpu().tolist()
    if rank == 0:
        steps = tot[2] / a.problems
        line = {'metric': METRIC, 'value': tot[1] * steps / mx[0] / 1e6, 'unit': UNIT, 'n_gpus': world,
                'steps': steps, 'warmup': 2, 'ms_per_step': mx[0] * 1e3 / max(steps, 1) / (a.problems / world),
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                'config': {'workload': 'cfg5: %d independent 7-camera x 5000-detection problems, RS + motion F, '
                                       'partitioned round-robin over %d GPU(s), %d in flight per GPU; end-to-end Scene.BA calls' % (a.problems, world, a.batch_threads)},
                'problems_per_s': a.problems / mx[0], 'device_ms_sum_max_rank': mx[3], 'gpu_launches': int(tot[4]),
                'e2e': {'value': tot[1] * steps / mx[0] / 1e6, 'unit': UNIT, 'seconds': mx[0]}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    global BA_KW
    wl_kw = {}
    if a.workload in ('cfg2', 'cfg3'):
        a.cams, a.det, a.coef = 7, 100000, 0
        wl_kw = dict(frames_per_knot=15.0)
        if a.workload == 'cfg3':
            wl_kw.update(opt_calib=True, motion_type='KE')
            BA_KW = dict(rs=True, motion_reg=True, motion_weights=1e2)
    if a.workload == 'cfg5' and a.impl == 'ours':
        return run_cfg5(a, rank, world, local)
    cfg = {'workload': workload_name(a), 'cams': a.cams, 'det_per_cam': a.det, 'coef_per_axis': a.coef,
           'l2_policy': 'inputs (J planes >= 0.3 GB) exceed L2; no flush needed',
           'parallelism': 'detections sharded along the global time axis (owner ranges of the sharded solve) over %d GPU(s); NCCL: camera blocks all-reduced, spline-side halo rows reduced to their owner' % world}

    if a.impl == 'reference':
        if rank != 0:
            return
        val, info, dt, iters = cpu_reference_run(a, a.steps, a.warmup)
        cfg['sample'] = info['sample']
        line = {'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': a.gpus, 'steps': iters, 'warmup': a.warmup,
                'ms_per_step': dt / iters * 1e3, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
                'dtype': 'f64', 'data': 'synthetic', 'config': cfg, 'impl': 'reference', 'cpu_baseline': info,
                'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                'gpu_launches': 0, 'lm_iters_per_s': info['lm_iters_per_s']}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from mvus_b200 import _cabi, ba, shard
    from mvus_b200.problem import FlatProblem
    torch.cuda.set_device(local)
    ba.DEVICE = local
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        shard.init_comm()

    t0 = time.perf_counter()
    fl = make_workload(a.cams, a.det, a.coef, **wl_kw)
    t_gen = time.perf_counter() - t0
    N_total = int(sum(d.shape[1] for d in fl.detections))
    if world > 1:
        # cut along the control-point ranges the ranks own in the sharded solve (halo-only exchange)
        fl = shard.shard_scene(fl, rank, world, shard.shard_bounds(fl, world, motion_reg=BA_KW.get('motion_reg', False)))
    fp = FlatProblem(fl, fl.numCam, **BA_KW)

    # ---- device-resident timing -----------------------------------------------------------
    hd = _cabi.Handle(fp, device=local, ftol=0.0, xtol=0.0, gtol=0.0, max_nfev=a.warmup + 1)
    if world > 1:
        hd.comm_init(*ba._COMM)
    hd.solve(fp.x0, want_r=False)                                   # warm-up steps
    ms_k1 = hd.time_resjac(fp.x0, reps=3)                           # K1+K1m alone (roofline)
    ms_k2 = hd.time_accumulate(reps=3)                              # K2+K2m alone
    hd.lib.mvus_ba_destroy(hd.h)
    hd.h = None
    hd = _cabi.Handle(fp, device=local, ftol=0.0, xtol=0.0, gtol=0.0, max_nfev=a.steps + 1)
    if world > 1:
        hd.comm_init(*ba._COMM)
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    x, _, st = hd.solve(fp.x0, want_r=False)
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    steps_done = st.nfev - 1
    ms = torch.tensor([st.ms_total, ms_k1, ms_k2, st.ms_resjac, st.ms_accum, st.ms_solve, st.ms_trial,
                       st.ms_syrk, st.ms_bcr, st.ms_reduce, st.ms_k2],
                      dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.cpu().tolist()
    hd.close()
    value = N_total * steps_done / (ms[0] / 1e3) / 1e6

    # ---- end to end through the public API (host buffers) -----------------------------------
    if world > 1:
        dist.barrier()
    # one untimed call first (warm-up: pinned output pool, CUDA context paths), then restore the
    # initial parameters so that the timed call does the same work from the same start
    ba.bundle_adjust(fl, fl.numCam, max_iter=2, ftol=0.0, xtol=0.0, gtol=0.0, **BA_KW)
    fp.unpack_into(fl, fp.x0)
    if world > 1:
        dist.barrier()
    # five timed repetitions from the same start; the MEDIAN is reported, all are listed (the host side of the
    # call -- page-locked output buffers, 7.6 GB of D2H -- is noisy: single repetitions 2x off were observed)
    import gc
    e2e_all, e2e_hosts, e2e_infos = [], [], []
    for _rep in range(5):
        fp.unpack_into(fl, fp.x0)
        gc.collect()                       # (the previous result's pinned arrays go back to the pool first)
        if world > 1:
            dist.barrier()
        pool0 = (_cabi.POOL.new_bytes, _cabi.POOL.reused_bytes)
        t0 = time.perf_counter()
        res_k = ba.bundle_adjust(fl, fl.numCam, max_iter=a.steps + 1, ftol=0.0, xtol=0.0, gtol=0.0, **BA_KW)
        torch.cuda.synchronize()
        dt_k = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(dt_k, op=dist.ReduceOp.MAX)
        e2e_all.append(float(dt_k.item()))
        e2e_hosts.append(res_k.stats.get('host'))
        e2e_infos.append({'nfev': int(res_k.nfev), 'host': res_k.stats.get('host'),
                          'pinned_new': _cabi.POOL.new_bytes - pool0[0], 'pinned_reused': _cabi.POOL.reused_bytes - pool0[1]})
        del res_k
    k_med = int(np.argsort(e2e_all)[len(e2e_all) // 2])
    e2e_info = e2e_infos[k_med]
    dt_e2e = e2e_all[k_med]
    e2e_steps = max(e2e_info['nfev'] - 1, 1)
    e2e_val = N_total * e2e_steps / dt_e2e / 1e6
    h2d = (3 * fp.N * 8 + fp.n * 8) / e2e_steps
    d2h = (fp.n * 8 + (2 * fp.N + hd.M) * 8 + 2 * 3 * fp.N * 8) / e2e_steps

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s'
    # FP64 tensor peak: MEASURED_PEAKS.json has no FP64 entry; measured on this pool's B200 with
    # profiles/micro/dmma_bench.cu (profiles/r2_fp64_bars.txt): 36.8 TFLOP/s for every mma.sync f64 shape
    fp64_peak = float(peaks.get('fp64_tensor_tflops', 36.8))
    fp64_src = ('measured (MEASURED_PEAKS.json)' if 'fp64_tensor_tflops' in peaks else
                'measured with profiles/micro/dmma_bench.cu (profiles/r2_fp64_bars.txt); MEASURED_PEAKS.json has no FP64 entry')
    P = fp.P
    N_loc = fp.N
    n_solves = max(int(st.lm_iterations), 1)
    ncP = fp.nc * fp.Pc
    rows_w = 3 * fp.n_ctrl / world                        # rows of W~ per rank
    k1_bytes = N_loc * (24 + 16 + 4 + 16 * P)            # 380 B/det (P=21)
    k2_bytes = N_loc * (16 + 4 + 16 * P)                 # 356 B/det
    bcr_bytes = 2.0 * rows_w * (ncP + 1) * 8             # one read + one write of W~ per solve
    syrk_flop = float(ncP) * (ncP + 1) * rows_w          # lower triangle of W~^T W~: n (n+1) / 2 x 2 flop x rows
    ms_syrk1, ms_bcr1 = ms[7] / n_solves, ms[8] / n_solves
    # DRAM traffic per launch from this round's `ncu --set full` captures (profiles/r2_ncu_traffic.json, written by
    # profiles/ncu_traffic.py from the .ncu-rep files: dram__bytes_read.sum + dram__bytes_write.sum), scaled by the
    # rank's share of the config-4 problem; None for other problem shapes
    traffic = {}
    try:
        tj = json.load(open(os.path.join(ROOT, 'profiles', 'r2_ncu_traffic.json')))
        if P == 21 and a.workload == 'cfg4':
            traffic = {k: v['dram_bytes'] * (N_loc / v['detections']) / 1e9 for k, v in tj.items()}
    except Exception:
        pass
    phases = {'resjac_K1': {'ms': ms[1], 'share_ms': ms[3], 'bound': 'hbm', 'achieved': k1_bytes / ms[1] / 1e6, 'unit': 'GB/s',
                            'frac': k1_bytes / ms[1] / 1e6 / peak, 'bytes_per_det': 44 + 16 * P, 'traffic': traffic.get('resjac_K1')},
              'accumulate_K2': {'ms': ms[2], 'share_ms': ms[10], 'bound': 'hbm', 'achieved': k2_bytes / ms[2] / 1e6, 'unit': 'GB/s',
                                'frac': k2_bytes / ms[2] / 1e6 / peak, 'bytes_per_det': 20 + 16 * P,
                                'traffic': traffic.get('accumulate_K2'),
                                'note': 'K2 + K2m + memsets of W~/D/E + chunk sort + band_to_blocks, run alone (3 reps)'},
              'schur_syrk_K3': {'ms': ms_syrk1, 'share_ms': ms[7], 'bound': 'tensor', 'achieved': syrk_flop / ms_syrk1 / 1e9 if ms_syrk1 else None,
                                'unit': 'TFLOP/s', 'frac': syrk_flop / ms_syrk1 / 1e9 / fp64_peak if ms_syrk1 else None,
                                'flop_per_launch': syrk_flop, 'traffic': traffic.get('schur_syrk_K3'),
                                'note': 'FP64 mma.sync (DMMA); useful flops = lower triangle n(n+1) x rows; per linear solve'},
              'cyclic_reduction_K3': {'ms': ms_bcr1, 'share_ms': ms[8], 'bound': 'hbm', 'achieved': bcr_bytes / ms_bcr1 / 1e6 if ms_bcr1 else None,
                                      'unit': 'GB/s', 'frac': bcr_bytes / ms_bcr1 / 1e6 / peak if ms_bcr1 else None,
                                      'bytes_per_launch': bcr_bytes, 'traffic': traffic.get('cyclic_reduction_K3'),
                                      'note': 'elimination of W~ in one solve: chunk pre-reduction (chunk_factor + chunk_w) + cyclic reduction of the chunk heads; algorithmic 2 |W~|; traffic = chunk_w_kernel alone'},
              'solve_share_ms': ms[5], 'resjac_share_ms': ms[3], 'accum_share_ms': ms[4], 'reduce_share_ms': ms[9],
              'trial_share_ms': ms[6], 'linear_solves': n_solves}
    kernels = ['resjac_K1', 'accumulate_K2', 'schur_syrk_K3', 'cyclic_reduction_K3']
    dom = max(kernels, key=lambda k: phases[k]['share_ms'])          # the kernel with the largest share of the timed region
    pd = phases[dom]
    roof = {'bound': pd['bound'], 'kernel': dom, 'achieved': pd['achieved'], 'peak': peak if pd['bound'] == 'hbm' else fp64_peak,
            'unit': pd['unit'], 'frac': pd['frac'], 'traffic': pd.get('traffic'),
            'traffic_unit': 'GB per launch (ncu dram__bytes_read+write of this round, profiles/r2_ncu_traffic.json)',
            'peak_source': peak_src if pd['bound'] == 'hbm' else fp64_src,
            'share_of_step': pd['share_ms'] / ms[0],
            'note': 'dominant kernel = largest share of the timed region (CUDA events inside the library)',
            'other_kernels': [{'kernel': k, 'bound': phases[k]['bound'], 'achieved': phases[k]['achieved'], 'unit': phases[k]['unit'],
                               'frac': phases[k]['frac'], 'share_of_step': phases[k]['share_ms'] / ms[0],
                               'traffic': phases[k].get('traffic')} for k in kernels if k != dom],
            'hbm_peak': peak, 'hbm_peak_source': peak_src, 'fp64_tensor_peak': fp64_peak}
    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': steps_done, 'warmup': a.warmup,
            'ms_per_step': ms[0] / max(steps_done, 1), 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': cfg,
            'lm_iters_per_s': steps_done / (ms[0] / 1e3),
            'resid_jac_mdet_per_s': N_total / (ms[1] / 1e3) / 1e6,
            'roofline': roof, 'phases': phases, 'clocks': clocks,
            'e2e': {'value': e2e_val, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'seconds': dt_e2e, 'seconds_all': e2e_all, 'reported': 'median of %d repetitions' % len(e2e_all), 'host_phases_all': e2e_hosts, 'steps': e2e_steps, 'host_phases_ms': e2e_info['host'],
                    'pinned_new_bytes': e2e_info['pinned_new'], 'pinned_reused_bytes': e2e_info['pinned_reused']},
            'gpu_launches': int(st.launches), 'linear_solves': int(st.lm_iterations), 'final_cost': st.cost, 'cost0': st.cost0,
            'workload_gen_s': t_gen}
    if world == 1 and not a.no_cpu_baseline:
        _, info, _, _ = cpu_reference_run(a, 9, 1)
        try:
            info['same_config'] = gpu_on_sample(a, 9)        # GPU and CPU on the SAME sample
            info['same_config']['cpu_value'] = info['value']
        except Exception as e:                               # (never lose the bench line over the side measurement)
            info['same_config'] = {'error': repr(e)}
        line['cpu_baseline'] = info
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
