#!/usr/bin/env python
"""Benchmark of the affinity U-Net watershed path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One step = one pass of the hot path over one synthetic 33x512x512 zyx frame
(BASELINE.json configs[1]): chunked U-Net (36 chunks of (10,256,256), margin
(1,64,64), all batched) -> seeds / Otsu mask / components -> exact flood.
With N > 1 (torchrun, one rank per GPU) every rank segments its own frame per
step (frame-wise sharding, weak scaling) and the ranks all-gather their label
counts over NCCL to make label ids global.

Printed (rank 0, one JSON line): `value` = voxels/s with the frame resident in
HBM, device-timed with CUDA events, max over ranks; `e2e` = the same through the
public frame loop `segmentation.segmentation_loop` (what `segment_data` runs) over a
pinned tzyx series (H2D of every frame, D2H of its labels inside the timed region);
`roofline` for the dominant kernel family (tcgen05 conv3d; tensor bound);
`cpu_baseline` = the oracle port of the reference timed on this box's host
cores on a bounded sample.  `--impl reference` times only that CPU path.
`--segmenter dog` measures the plugin's second segmenter (configs[4]) with the same contract.
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

FRAME = (33, 512, 512)
CHUNK = (10, 256, 256)
MARGIN = (1, 64, 64)
METRIC = 'voxels/sec end-to-end affinity U-Net watershed'
WORKLOAD = ('configs[1]: one synthetic platelet frame 33x512x512 zyx per step, chunk (10,256,256) '
            'margin (1,64,64) = 36 chunks batched, fp16 tcgen05 U-Net + GPU seeds/mask/CCL/flood')


def measured_peaks():
    fn = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(fn):
        with open(fn) as f:
            p = json.load(f)
        return p, 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


def conv_traffic():
    """DRAM bytes per conv3d_tc launch (average over the 16 layers) from the committed
    `ncu --set full` capture of the same kernels (profiles/r01_conv_traffic.json), else None."""
    fn = os.path.join(ROOT, 'profiles', 'r01_conv_traffic.json')
    try:
        with open(fn) as f:
            return float(json.load(f)['dram_bytes_per_launch'])
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark(self):
        """Start of the timed region: the sampler itself is started before the warm-up so that
        nvidia-smi is up by then; only samples taken after mark() count."""
        self.t_mark = time.time()

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        t_mark = getattr(self, 't_mark', 0.0)
        rows = [r for ts, r in self.rows if ts >= t_mark]
        window = 'timed region'
        if not rows and self.rows:            # a very short timed region: the last samples under load
            rows = [r for ts, r in self.rows[-3:]]
            window = 'warm-up (timed region shorter than the sampling period)'
        self.window = window
        for r in rows:
            f = [x.strip() for x in r.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': float(max(smax)) if smax else None,
                'reasons': sorted(reasons), 'samples': len(sm), 'window': window}


# --------------------------------------------------------------------------------------
# CPU reference arm (the oracle port of the reference, oracle/*)
# --------------------------------------------------------------------------------------
def cpu_reference_step(vol, labels_gt, sd, n_sample_chunks, threads):
    """One bounded sample of the reference CPU path on this frame: `n_sample_chunks` of the
    36 U-Net chunks (fp32 torch/oneDNN, train-mode BN, as predict.py:100-126 with the U-Net
    on the CPU) + the whole post-U-Net stage (watershed.py:165-223) on analytic feature maps
    of the same frame.  Returns the extrapolated seconds per frame and the split."""
    import torch
    from oracle import chunks as ochunks, post as opost, unet_ref
    from iterseg_b200 import synth
    torch.set_num_threads(threads)
    starts, _ = ochunks.make_chunks(vol.shape, CHUNK, MARGIN)
    pick = [starts[i] for i in np.linspace(0, len(starts) - 1, n_sample_chunks).astype(int)]
    t0 = time.perf_counter()
    for st in pick:
        sl = tuple(slice(s, s + c) for s, c in zip(st, CHUNK))
        x = torch.from_numpy(np.ascontiguousarray(vol[sl])[None, None])
        unet_ref.unet_forward(x, sd)
    t_unet = (time.perf_counter() - t0) / len(pick)
    feats = synth.analytic_features(labels_gt, 0)
    out = np.zeros(tuple(s + 2 for s in vol.shape), np.uint32)
    t0 = time.perf_counter()
    opost.segment_output_image(feats, out=out.ravel())
    t_post = time.perf_counter() - t0
    per_frame = t_unet * len(starts) + t_post
    return per_frame, t_unet, t_post, len(starts)


def run_reference(args, rank):
    if rank != 0:
        return 0
    from iterseg_b200 import synth
    from oracle import flood as oflood
    oflood.build()
    threads = os.cpu_count() or 1
    vol, lab = synth.platelet_frame(FRAME, seed=0, return_labels=True)
    sd = synth.structured_state_dict(0)
    n_sample = 2
    for _ in range(args.warmup):
        cpu_reference_step(vol, lab, sd, 1, threads)
    times, splits = [], []
    for _ in range(args.steps):
        per_frame, t_unet, t_post, n_chunks = cpu_reference_step(vol, lab, sd, n_sample, threads)
        times.append(per_frame)
        splits.append((t_unet, t_post))
    per_frame = float(np.mean(times))
    nvox = float(np.prod(FRAME))
    value = nvox / per_frame
    sample = (f'{n_sample} of {n_chunks} U-Net chunks per step timed (fp32 torch CPU, train-mode BN) and '
              f'extrapolated x{n_chunks}/{n_sample}; full post-U-Net stage (scipy/numpy + C heap flood, '
              f'single thread like the numba original) on analytic feature maps of the same frame; '
              f'mean U-Net {np.mean([s[0] for s in splits]):.2f} s/chunk, post {np.mean([s[1] for s in splits]):.2f} s/frame')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'voxels/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': per_frame * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'frame': list(FRAME), 'chunk': list(CHUNK), 'margin': list(MARGIN)},
        'cpu_baseline': {'value': value, 'unit': 'voxels/s', 'cores': threads, 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': value, 'unit': 'voxels/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
def run_gpu(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from iterseg_b200 import _lib, distributed as idist, predict, segmentation, synth, unet as unet_mod
    from iterseg_b200 import watershed as ws
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    lib = _lib.load()
    _lib.require_device()
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- assets (untimed): synthetic frame of this rank, synthetic network file -----------
    vol_np, lab_gt = synth.platelet_frame(FRAME, seed=rank, return_labels=True)
    sd = synth.structured_state_dict(0)
    net = unet_mod.UNet()
    net.load_state_dict(sd)
    net.to(dev)
    frame = torch.from_numpy(vol_np).to(dev)
    shape_p = tuple(s + 2 for s in FRAME)
    labels = torch.zeros(shape_p, dtype=torch.int32, device=dev)
    feats = torch.zeros((5,) + FRAME, dtype=torch.float32, device=dev)
    nvox = float(np.prod(FRAME))

    from iterseg_b200.pipeline import FramePipeline
    pipe = FramePipeline(net, FRAME, CHUNK, MARGIN)

    def steps_device(k):
        """k complete frames (all kernels of every frame inside the call): the post stage of
        frame i overlaps the U-Net of frame i+1 on a second stream (iterseg_b200/pipeline.py)."""
        counts = None
        pipe.submit(frame)
        for i in range(k):
            if i + 1 < k:
                pipe.submit(frame)
            lab, counts = pipe.collect()
            if world > 1:
                with torch.cuda.stream(pipe.s_post):
                    n_local = int(counts[0].item())
                    all_counts = idist.gather_label_counts({rank: n_local}, world, rank, world, device=dev)
                    idist.add_label_offset_(lab, int(idist.exclusive_offsets(all_counts)[rank]))
        pipe.drain_to()
        return counts

    # pinned host buffers for the end-to-end (public API) measurement: a K-frame tzyx series
    # through `segmentation.segmentation_loop`, the frame loop behind `segment_data` /
    # `affinity_unet_watershed` (labels restart at 1 in every frame, as in the reference)
    n_e2e = max(args.steps, 2)
    series = torch.from_numpy(np.broadcast_to(vol_np, (n_e2e,) + FRAME).copy()).pin_memory()
    out_series = torch.zeros((n_e2e,) + FRAME, dtype=torch.int32).pin_memory()
    config = {'unet': net, 'output_volume': np.zeros((1,), np.float32)}

    def run_e2e():
        out = out_series.numpy()
        out[...] = 0                       # untimed: the caller's fresh output store
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        done = list(segmentation.segmentation_loop(None, series.numpy(), CHUNK, MARGIN, out,
                                                   segmentation.affinity_watershed_for_chunks, config))
        torch.cuda.synchronize()
        assert len(done) == n_e2e
        return time.perf_counter() - t0

    sampler = ClockSampler(local_rank)
    sampler.start()                       # before the warm-up: nvidia-smi needs ~0.2 s to come up
    counts = steps_device(max(args.warmup, 1))
    barrier()
    plan = list(net._plans.values())[0]
    _lib.check(lib.isg_unet_plan_profile(plan.ptr, 1), 'profile')
    launches0 = lib.isg_launch_count()
    sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    counts = steps_device(args.steps)
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = int(lib.isg_launch_count() - launches0)
    ms_total = e0.elapsed_time(e1)
    prof = (ctypes_double_array(5))
    _lib.check(lib.isg_unet_plan_profile_read(plan.ptr, prof), 'profile_read')
    _lib.check(lib.isg_unet_plan_profile(plan.ptr, 0), 'profile')
    tc_ms, n_tc, fw_ms, n_fw, tc_flops = [float(x) for x in prof]
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = nvox * world / (ms_step * 1e-3)

    # ---- end to end through the public frame loop, host buffers ---------------------------------
    run_e2e()                                   # warm-up
    e2e_runs = []
    for _ in range(3):                          # host wall clock is noisy on a shared box: median of 3
        barrier()
        e2e_runs.append(run_e2e())
    dt = sorted(e2e_runs)[1]
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = nvox * world * n_e2e / float(t.item())
    e2e_labels_ok = bool(int(out_series[0].max()) == int(counts[0].item()) or world > 1)

    # ---- post stage alone (not overlapped), device-timed: the HBM-side roofline entry -----------
    post_ms = None
    if rank == 0:
        predict.predict_frame_device(net, frame, CHUNK, MARGIN, out=feats)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        reps = 3
        p0.record()
        for _ in range(reps):
            labels.zero_()
            ws.segment_features_device(feats, labels)
        p1.record()
        torch.cuda.synchronize()
        post_ms = p0.elapsed_time(p1) / reps
    if rank == 0:
        peaks, peak_kind = measured_peaks()
        peak_tf = float(peaks.get('bf16_tflops_sustained', peaks.get('bf16_tflops')))
        ach_tf = (tc_flops * n_fw) / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
        counts_h = [int(x) for x in counts.cpu().numpy()]
        line = {
            'metric': METRIC, 'value': value, 'unit': 'voxels/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f16',
            'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'frame': list(FRAME), 'chunk': list(CHUNK),
                       'margin': list(MARGIN), 'frames_per_step': world,
                       'parallelism': f'frames x{world}' if world > 1 else 'single GPU',
                       'pipeline': 'post stage of frame i overlaps the U-Net of frame i+1 (two streams); '
                                   'every frame of the timed region is completed inside it',
                       'network': 'synthetic state_dict (structured carriers + dense random weights), '
                                  'fp16 operands / fp32 accumulate (bf16 misses the 1e-2 parity gate)',
                       'l2': 'inputs larger than L2: ~13.8 GB of activations streamed per step',
                       'objects': {'seeds': counts_h[0], 'components': counts_h[2],
                                   'multi_seed_components': counts_h[3]}},
            'e2e': {'value': e2e_value, 'unit': 'voxels/s',
                    'h2d_bytes_per_step': int(np.prod(FRAME) * 4) * world,
                    'd2h_bytes_per_step': int(np.prod(FRAME) * 4) * world,
                    'api': f'segmentation.segmentation_loop over a pinned {n_e2e}-frame tzyx series (the frame loop '
                           f'of segment_data): per frame H2D, min/max + normalise, U-Net, post stage, D2H into the '
                           f'caller\'s int32 array; two frames in flight; median of 3 timed runs',
                    'runs_s': [round(x, 5) for x in e2e_runs], 'labels_match_device_run': e2e_labels_ok},
            'gpu_launches': launches,
            'clocks': clocks,
            'roofline': {'bound': 'tensor', 'achieved': ach_tf, 'peak': peak_tf, 'unit': 'TFLOP/s',
                         'frac': ach_tf / peak_tf if peak_tf else None, 'traffic': conv_traffic(),
                         'kernel': 'conv3d_tc_kernel x13 + conv3d_zring32_kernel x2 + conv3d_zring_kernel (the 16 TMA-fed tcgen05 implicit-GEMM launches per step)',
                         'peak_source': f'{peak_kind} bf16_tflops_sustained (kernel timed inside a long step)',
                         'share_of_step': (tc_ms / n_fw) / ms_step if n_fw else None,
                         'unet_ms_per_step': fw_ms / n_fw if n_fw else None},
        }
        hbm = float(peaks.get('hbm_gbs_sustained', peaks.get('hbm_gbs', 6650.0)))
        post_bytes = 24.0 * nvox              # 5 x f32 feature reads + 1 x u32 label write per voxel
        line['roofline_post'] = {
            'bound': 'hbm', 'achieved': post_bytes / (post_ms * 1e-3) / 1e9, 'peak': hbm, 'unit': 'GB/s',
            'frac': post_bytes / (post_ms * 1e-3) / 1e9 / hbm, 'traffic': None,
            'kernel': 'post-U-Net stage (seeds, Otsu mask, components, ordered flood), timed alone',
            'ms': post_ms,
            'note': 'latency bound, not HBM bound: the order-exact flood of the largest multi-seed object '
                    '(one warp, ~660 clk per voxel) sets the time; see profiles/r01_notes.md'}
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            per_frame, t_unet, t_post, n_chunks = cpu_reference_step(vol_np, lab_gt, sd, 3, threads)
            line['cpu_baseline'] = {
                'value': nvox / per_frame, 'unit': 'voxels/s', 'cores': threads, 'kind': 'port',
                'sample': (f'3 of {n_chunks} U-Net chunks timed on {threads} threads ({t_unet:.2f} s/chunk, fp32 '
                           f'torch CPU, train-mode BN) extrapolated to {n_chunks}; full post-U-Net stage '
                           f'({t_post:.2f} s, single thread) on analytic feature maps of the same frame')}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


# --------------------------------------------------------------------------------------
# DoG blob segmenter arm (BASELINE.json configs[4]):  bench.py --segmenter dog
# --------------------------------------------------------------------------------------
def run_dog(args, rank, local_rank, world):
    """One step = the DoG blob watershed (min_sigma 1, max_sigma 1.5, threshold 0.02) of one
    synthetic 33x512x512 frame per rank.  Same JSON contract; `cpu_baseline` is the scipy + C
    restatement (oracle/dog.py) on the same frame, one thread."""
    import torch
    import torch.distributed as dist
    from iterseg_b200 import _lib, segmentation, synth
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    lib = _lib.load()
    _lib.require_device()
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    vol = synth.platelet_frame(FRAME, seed=rank)
    frame = torch.from_numpy(vol).to(dev)
    shape_p = tuple(s + 2 for s in FRAME)
    labels = torch.zeros(shape_p, dtype=torch.int32, device=dev)
    cfg = dict(min_sigma=1, max_sigma=1.5, threshold=0.02)
    nvox = float(np.prod(FRAME))

    def step():
        labels.zero_()
        return segmentation.dog_blob_segment_device(frame, labels, **cfg)

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        mask, counts = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = lib.isg_launch_count()
    sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        mask, counts = step()
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    launches = int(lib.isg_launch_count() - launches0)
    t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
    vol_pinned = torch.from_numpy(vol.copy()).pin_memory()
    out_pinned = torch.zeros(shape_p, dtype=torch.int32).pin_memory()

    def step_e2e():
        segmentation.dog_blob_watershed_for_chunks(vol_pinned.numpy(), out_pinned.numpy().view(np.uint32),
                                                   CHUNK, MARGIN, **cfg)

    for _ in range(2):
        step_e2e()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    te = torch.tensor([(time.perf_counter() - t0) / args.steps * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    if rank == 0:
        peaks, peak_kind = measured_peaks()
        hbm = float(peaks.get('hbm_gbs_sustained', peaks.get('hbm_gbs', 6650.0)))
        ms = float(t.item())
        c = counts.cpu().numpy()
        line = {
            'metric': 'voxels/sec DoG blob watershed', 'value': nvox * world / (ms * 1e-3), 'unit': 'voxels/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'configs[4]: DoG blob watershed (min_sigma 1, max_sigma 1.5, threshold 0.02), one '
                                   'synthetic platelet frame 33x512x512 per rank and step', 'frame': list(FRAME),
                       'parallelism': f'frames x{world}' if world > 1 else 'single GPU',
                       'objects': {'blobs': int(c[1]), 'labels': int(labels.max().item())}},
            'e2e': {'value': nvox * world / (float(te.item()) * 1e-3), 'unit': 'voxels/s',
                    'h2d_bytes_per_step': int(vol_pinned.numel() * 4) * world,
                    'd2h_bytes_per_step': int(out_pinned.numel() * 4) * world},
            'gpu_launches': launches, 'clocks': clocks,
            'roofline': {'bound': 'hbm', 'achieved': 8.0 * nvox / (ms * 1e-3) / 1e9, 'peak': hbm, 'unit': 'GB/s',
                         'frac': 8.0 * nvox / (ms * 1e-3) / 1e9 / hbm, 'traffic': None,
                         'kernel': 'whole DoG stage (12 separable Gaussian passes, peaks, EDT, flood)',
                         'peak_source': f'{peak_kind} hbm_gbs; 8 B per voxel compulsory (f32 in, i32 out)'},
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import dog
            out = np.zeros(shape_p, np.int32)
            t0 = time.perf_counter()
            dog.dog_blob_watershed_for_chunks(vol, out, **cfg)
            dt = time.perf_counter() - t0
            line['cpu_baseline'] = {'value': nvox / dt, 'unit': 'voxels/s', 'cores': 1, 'kind': 'port',
                                    'sample': 'one whole frame: scipy.ndimage Gaussians / EDT / label + C heap flood'}
            line['identical_to_cpu_restatement'] = bool(np.array_equal(out, labels.cpu().numpy()))
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def ctypes_double_array(n):
    import ctypes
    return (ctypes.c_double * n)()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--segmenter', default='affinity', choices=['affinity', 'dog'],
                    help="'dog': the DoG blob watershed (BASELINE.json configs[4]) instead of the headline path")
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        return run_reference(args, rank)
    if args.warmup < 3:
        args.warmup = 3
    if args.segmenter == 'dog':
        return run_dog(args, rank, local_rank, world)
    return run_gpu(args, rank, local_rank, world)


if __name__ == '__main__':
    sys.exit(main())
