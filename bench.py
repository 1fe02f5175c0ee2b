#!/usr/bin/env python
"""bench.py -- headline benchmark of the DynaMask per-instance mask hot path on B200.

Workload (BASELINE.json configs[1], "C2"): RoIAlign fwd/bwd microbench, 512 RoIs/img x 256 ch x
4 FPN levels (800x1344 image), mixed 14/28/56/112 outputs, batch 16 per GPU.  One *step* is one
pass of the hot path over that batch: dm_assign -> dm_roi_align_fwd (all buckets, one launch) ->
dm_roi_align_bwd (zero-init of the gradient pyramid + one launch).  `value` is RoIs/s with all
inputs resident in HBM; `e2e` is the same pass through the plugin surface with HOST buffers
(pinned host -> device copies of the pyramid / RoIs / labels and device -> host copies of the
pooled features and the gradient pyramid inside the timed region).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU); images are sharded by rank, no collective on
the data path (NCCL only reduces timings / all-gathers checksums).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import synth  # noqa: E402

IMG_H, IMG_W = 800, 1344
STRIDES = [4, 8, 16, 32]
BUCKET_SIZES = (14, 28, 56, 112)
METRIC = 'rois_per_sec_mask_extract_fwd_bwd'
UNIT = 'RoIs/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--batch', type=int, default=16, help='images per GPU')
    ap.add_argument('--rois-per-img', type=int, default=512)
    ap.add_argument('--channels', type=int, default=256)
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true')
    ap.add_argument('--no-configs', action='store_true', help='skip the C3 / C4 / C5 workloads')
    ap.add_argument('--no-competitor', action='store_true', help='skip the torchvision-CUDA leg (e.g. under ncu)')
    ap.add_argument('--e2e-steps', type=int, default=2)
    ap.add_argument('--e2e-chunk-images', type=int, default=2)
    ap.add_argument('--cpu-sample-rois', type=int, default=512, help='RoIs per image in the CPU sample (config: 512)')
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` on the C2 workload, from the
    committed ncu capture (profiles/roofline_traffic.json); None when no capture is recorded."""
    p = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    try:
        return float(json.load(open(p))[kernel]['dram_bytes_per_launch'])
    except Exception:
        return None


# ----------------------------------------------------------------------------------------------
# algorithmic bytes
# ----------------------------------------------------------------------------------------------
def footprint_pixels(rois, lvl, shapes):
    """fp_r: feature pixels in each RoI's bilinear footprint at its level, clipped to the map."""
    r = rois.double()
    fp = torch.zeros(r.size(0), dtype=torch.float64)
    for l, (h, w) in enumerate(shapes):
        s = 1.0 / STRIDES[l]
        m = lvl == l
        if not bool(m.any()):
            continue
        x_lo = torch.clamp(torch.floor(r[m, 1] * s - 0.5), min=0)
        x_hi = torch.clamp(torch.floor(r[m, 3] * s - 0.5) + 1, max=w - 1)
        y_lo = torch.clamp(torch.floor(r[m, 2] * s - 0.5), min=0)
        y_hi = torch.clamp(torch.floor(r[m, 4] * s - 0.5) + 1, max=h - 1)
        fp[m] = torch.clamp(x_hi - x_lo + 1, min=0) * torch.clamp(y_hi - y_lo + 1, min=0)
    return fp


def algorithmic_bytes(rois, lvl, bucket, shapes, batch, channels):
    p2 = torch.tensor([s * s for s in BUCKET_SIZES], dtype=torch.float64)[bucket]
    fp = footprint_pixels(rois, lvl, shapes)
    k = rois.size(0)
    fwd = float((4.0 * channels * (p2 + fp)).sum()) + 20.0 * k
    pyramid = 4.0 * batch * channels * sum(h * w for h, w in shapes)
    bwd = float((4.0 * channels * (p2 + 2 * fp)).sum()) + pyramid
    return fwd, bwd


# ----------------------------------------------------------------------------------------------
# host placement for the e2e leg
# ----------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(index):
    """Run this rank on the CPUs NVML lists as local to its GPU, so that the pinned staging buffers
    of the e2e leg are allocated on (first touched from) the GPU's own NUMA node: with several ranks
    on one host the device<->host copies otherwise cross the socket interconnect.  Returns the
    number of CPUs bound to, or None when nothing was changed."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get('CUDA_VISIBLE_DEVICES', '')
        phys = index
        if vis and all(v.strip().isdigit() for v in vis.split(',')) and index < len(vis.split(',')):
            phys = int(vis.split(',')[index])
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return None


# ----------------------------------------------------------------------------------------------
# clocks sampler
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed region: NVML polled every few ms
    from a thread (the timed region of the default run is ~160 ms -- `nvidia-smi -lms` does not even
    deliver its first line in that time); nvidia-smi is the fall-back when pynvml is missing."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.nvml = None
        self.samples = []       # (sm_mhz, power_w, reasons bitmask)
        self.stop_flag = False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get('CUDA_VISIBLE_DEVICES', '')
            phys = self.index
            if vis and all(v.strip().isdigit() for v in vis.split(',')) and self.index < len(vis.split(',')):
                phys = int(vis.split(',')[self.index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    watts = nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                except Exception:
                    watts = 0.0
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((mhz, watts, mask))
            except Exception:
                pass
            time.sleep(0.004)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            nv = self.nvml
            if not self.samples:
                return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
            bits = {'hw_slowdown': nv.nvmlClocksThrottleReasonHwSlowdown,
                    'hw_thermal_slowdown': nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                    'sw_thermal_slowdown': nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                    'sw_power_cap': nv.nvmlClocksThrottleReasonSwPowerCap}
            seen = 0
            for _, _, m in self.samples:
                seen |= m
            return {'sm_mhz': float(np.median([x[0] for x in self.samples])), 'sm_max_mhz': self.max_mhz,
                    'power_w_max': float(max(x[1] for x in self.samples)), 'samples': len(self.samples),
                    'source': 'nvml', 'reasons': sorted(k for k, b in bits.items() if seen & b)}
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)),
                'power_w_max': float(max(power)), 'samples': len(sm), 'source': 'nvidia-smi',
                'reasons': sorted(reasons)}


# ----------------------------------------------------------------------------------------------
# CPU path (oracle port) used by cpu_baseline and --impl reference
# ----------------------------------------------------------------------------------------------
def _reference_extractors():
    """The reference's own, unmodified SingleRoIExtractor (one per pooled size) when its tree is
    present (the build container; /root/reference does not exist on the GPU box), else None.  The
    RoIAlign layer under it is torchvision's CPU kernel either way (mmcv is not installable)."""
    try:
        from oracle import ref_shim
        if not ref_shim.available():
            return None
        ns = ref_shim.load()
        return [ns.SingleRoIExtractor(dict(type='RoIAlign', output_size=p, sampling_ratio=0), 256, STRIDES)
                for p in BUCKET_SIZES]
    except Exception:
        return None


def cpu_extract_fwd_bwd(n_images, rois_per_img, channels, seed, threads):
    """Reference host path on CPU for a sample of C2: per image, the per-level select / RoIAlign /
    scatter loop at each RoI's selected size, plus autograd backward, with torchvision's CPU roi_align
    as the stand-in for the absent mmcv kernel.  The loop is the reference's unmodified
    SingleRoIExtractor where its tree exists, the oracle's restatement of it elsewhere.  Images are
    sharded over a thread pool (ATen releases the GIL).  Returns (rois, seconds, kind)."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import oracle as O
    g = torch.Generator().manual_seed(seed)
    work = []
    for _ in range(n_images):
        feats = synth.make_features(1, channels, IMG_H, IMG_W, g)
        rois = synth.make_rois(1, rois_per_img, IMG_H, IMG_W, g)
        onehot = synth.make_onehot(rois_per_img, g)
        work.append((feats, rois, onehot))
    ref_ext = _reference_extractors()

    def one(item):
        torch.set_num_threads(1)
        feats, rois, onehot = item
        fr = [f.clone().requires_grad_() for f in feats]
        if ref_ext is not None:
            bucket = onehot.argmax(1)
            outs = [ref_ext[b](fr, rois[bucket == b]) for b in range(len(BUCKET_SIZES)) if bool((bucket == b).any())]
        else:
            outs, _, _ = O.bucketed_extract(fr, rois, onehot, BUCKET_SIZES, STRIDES, kernel=O.roi_align_tv)
        loss = sum((o * o).sum() for o in outs) * 0.5   # grad_out = out, like the GPU step
        loss.backward()
        return float(fr[0].grad.abs().sum())

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(one, work))
    dt = time.perf_counter() - t0
    return n_images * rois_per_img, dt, ('reference' if ref_ext is not None else 'port')


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the
    reference is Python + un-vendored mmcv, nothing compiles into oracle/_ref)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 32))
    # one step = the config's batch: 16 images x 512 RoIs, one image per worker thread
    n_img = args.batch
    times, nrois, kind = [], 0, 'port'
    warm = args.warmup
    for i in range(args.warmup + args.steps):
        n, dt, kind = cpu_extract_fwd_bwd(n_img, args.cpu_sample_rois, args.channels, 1234 + i, threads)
        if i == 0 and dt > 20.0:
            warm = min(warm, 1)          # a step of many seconds needs no three warm-ups
        if i >= warm:
            times.append(dt)
            nrois = n
        if sum(times) > 120 or len(times) >= args.steps:
            break
    steps_done = max(len(times), 1)
    ms = 1000.0 * sum(times) / steps_done
    val = nrois / (ms / 1000.0)
    sample = ('%d images x %d RoIs (uniform 14/28/56/112 mix), C=%d, fwd+bwd, thread pool over images (%d workers); %s' % (
        n_img, args.cpu_sample_rois, args.channels, min(threads, n_img),
        "the reference's unmodified SingleRoIExtractor over torchvision's CPU roi_align" if kind == 'reference'
        else "oracle restatement of the reference loop over torchvision's CPU roi_align"))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': steps_done, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args),
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': min(threads, n_img), 'kind': kind, 'sample': sample},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def workload_config(args):
    return {
        'workload': 'C2 RoIAlign fwd+bwd microbench: %d RoIs/img x %d ch x 4 FPN levels (800x1344), '
                    'uniform 14/28/56/112 output mix, batch %d per GPU' % (
                        args.rois_per_img, args.channels, args.batch),
        'rois_per_step_per_gpu': args.rois_per_img * args.batch,
        'layout': 'NCHW fp32 in, NCHW fp32 out (reference layout)',
        'l2': 'no flush needed: each step streams ~75 GB through a 126 MB L2',
        'sharding': 'by image, no data-path collective',
    }


# ----------------------------------------------------------------------------------------------
# native arm
# ----------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import torch.distributed as dist

    import dynamask_b200 as dm
    from dynamask_b200 import _lib, ops

    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    B, C, R = args.batch, args.channels, args.rois_per_img
    shapes = synth.pyramid_shapes(IMG_H, IMG_W)
    g = torch.Generator().manual_seed(1234 + rank)
    gd = torch.Generator(device=dev).manual_seed(1234 + rank)
    feats = [torch.randn(B, C, h, w, generator=gd, device=dev) for (h, w) in shapes]
    rois_h = synth.make_rois(B, R, IMG_H, IMG_W, g)
    onehot_h = synth.make_onehot(rois_h.size(0), g)
    rois, onehot = rois_h.to(dev), onehot_h.to(dev)
    K = rois.size(0)
    scales = [1.0 / s for s in STRIDES]
    out_hw = [v for s in BUCKET_SIZES for v in (s, s)]
    feat_shapes = [int(v) for f in feats for v in f.shape]

    # bucket counts are host ints known before the loop (one readback at setup)
    lvl0, bucket0, _, seg0 = ops.assign(rois, onehot, 4, 56.0, 4)
    seg_h = seg0.cpu()
    counts = (seg_h[1:] - seg_h[:-1]).tolist()
    fwd_bytes, bwd_bytes = algorithmic_bytes(rois_h, lvl0.cpu().long(), bucket0.cpu().long(), shapes, B, C)

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    marks = []

    def step(record):
        e0, e1, e2, e3 = (ev(), ev(), ev(), ev()) if record else (None, ) * 4
        if record:
            e0.record()
        lvl, _, perm, seg = ops.assign(rois, onehot, 4, 56.0, 4)
        if record:
            e1.record()
        outs = ops.roi_align_forward(feats, rois, lvl, perm, seg, counts, out_hw, scales, 0, True, False)
        if record:
            e2.record()
        grads = ops.roi_align_backward(outs, rois, lvl, perm, seg, feat_shapes, [False] * 4, out_hw,
                                       scales, 0, True)
        if record:
            e3.record()
            marks.append((e0, e1, e2, e3))
        return outs, grads

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        outs, grads = step(False)
    del outs, grads
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = _lib.launch_count()
    t_start, t_end = ev(), ev()
    barrier()
    t_start.record()
    for _ in range(args.steps):
        outs, grads = step(True)
    t_end.record()
    barrier()
    clocks = sampler.stop()
    launches = _lib.launch_count() - launches0
    total_ms = t_start.elapsed_time(t_end)
    checksum = float(sum(float(gr.double().sum()) for gr in grads))
    del outs, grads

    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sums = [None] * world
        dist.all_gather_object(sums, checksum)
    else:
        sums = [checksum]
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * K / (ms_per_step / 1000.0)

    asg_ms = float(np.mean([a.elapsed_time(b) for a, b, _, _ in marks]))
    fwd_ms = float(np.mean([b.elapsed_time(c) for _, b, c, _ in marks]))
    bwd_ms = float(np.mean([c.elapsed_time(d) for _, _, c, d in marks]))
    peak, peak_src = peaks()
    kern = {
        'dm_roi_align_fwd': {'ms': fwd_ms, 'algorithmic_bytes': fwd_bytes,
                             'achieved_gbs': fwd_bytes / fwd_ms / 1e6, 'rois_per_s': K / fwd_ms * 1e3},
        'dm_roi_align_bwd(+zero-init)': {'ms': bwd_ms, 'algorithmic_bytes': bwd_bytes,
                                        'achieved_gbs': bwd_bytes / bwd_ms / 1e6,
                                        'rois_per_s': K / bwd_ms * 1e3},
        'dm_assign': {'ms': asg_ms},
    }
    for v in kern.values():
        if 'achieved_gbs' in v:
            v['frac_of_measured_peak'] = v['achieved_gbs'] / peak
            v['frac_of_8TBs_spec'] = v['achieved_gbs'] / 8000.0
    dom = 'dm_roi_align_bwd(+zero-init)' if bwd_ms >= fwd_ms else 'dm_roi_align_fwd'
    roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': kern[dom]['achieved_gbs'], 'peak': peak,
                'peak_source': peak_src, 'unit': 'GB/s', 'frac': kern[dom]['achieved_gbs'] / peak,
                'traffic': ncu_traffic(dom)}

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(args),
        'roofline': roofline, 'kernels': kern, 'gpu_launches': int(launches), 'clocks': clocks,
        'checksums': sums,
    }

    # which of the absent third-party libraries (the arithmetic of the RLE / polygon / SimpleRoIAlign rows lives
    # in them) could be imported here: when False the rows' oracles are restatements, "parity unpinned"
    pins = {}
    for name in ('pycocotools', 'mmcv'):
        try:
            pins[name] = getattr(__import__(name), '__file__', None) is not None
        except Exception:
            pins[name] = False
    line['third_party_importable'] = pins

    # ---- extras: the other two kernels on their own configs (rank 0 reports) ------------------
    if not args.no_extras:
        line['extras'] = run_extras(dm, ops, dev, rank, peak)

    # ---- BASELINE.json configs[2..4] on EVERY rank (times reduced with MAX, units summed) ---------
    if not args.no_configs:
        import bench_configs
        line['configs'] = bench_configs.run_all(dm, ops, dev, rank, world, peak)

    if not args.no_extras and not args.no_competitor and rank == 0 and world == 1:
        line['gpu_competitor'] = run_competitor(dm, dev, rank, feats, rois, onehot)
        if 'c2_torchvision_cuda_ms' in line['gpu_competitor']:
            line['gpu_competitor']['c2_ours_ms'] = fwd_ms + bwd_ms + asg_ms

    # ---- e2e: plugin surface with host buffers -------------------------------------------------
    if not args.no_e2e:
        # the e2e leg runs on the CPUs local to this rank's GPU (pinned staging buffers on the GPU's NUMA
        # node); the affinity is restored afterwards so that the CPU baseline leg sees every core
        aff0 = os.sched_getaffinity(0)
        numa = bind_to_gpu_numa_node(local)
        e2e = run_e2e(args, dm, dev, rank, world, feats, rois_h, onehot_h, counts)
        os.sched_setaffinity(0, aff0)
        line['e2e'] = e2e
        if numa:
            line['e2e']['cpus_bound_per_rank'] = numa
        line['e2e'].update(host_copy_ceiling(world, e2e))
    del feats
    torch.cuda.empty_cache()

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        threads = max(1, min(cores, 32))
        # bounded sample: half the config's batch at the config's 512 RoIs per image (one image per worker)
        n_img = max(1, min(threads, args.batch // 2))
        n, dt, kind = cpu_extract_fwd_bwd(n_img, args.cpu_sample_rois, C, 1234, threads)
        line['cpu_baseline'] = {
            'value': n / dt, 'unit': UNIT, 'cores': min(threads, n_img), 'kind': kind,
            'sample': '%d images x %d RoIs of the C2 workload (uniform size mix), fwd+bwd, one pass, '
                      'torchvision CPU roi_align under the %s, thread pool over images' % (
                          n_img, args.cpu_sample_rois,
                          "reference's unmodified SingleRoIExtractor" if kind == 'reference' else 'reference host loop (oracle port)')}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_extras(dm, ops, dev, rank, peak):
    ex = {}
    g = torch.Generator().manual_seed(99 + rank)
    # paste: C4 shape per GPU, 8 images x 100 detections pasted into 800x1333 canvases, one launch
    n = 800
    logits = synth.make_mask_logits(n, 112, g).to(dev)
    boxes = synth.make_boxes(n, 800, 1333, g, s_lo=8, s_hi=500).to(dev)
    for _ in range(3):
        out = ops.paste_masks(logits, boxes, None, 800, 1333, [0, 0, 1333, 800], True, 0.5, ops.PASTE_BOOL)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    reps = 10
    for _ in range(reps):
        out = ops.paste_masks(logits, boxes, None, 800, 1333, [0, 0, 1333, 800], True, 0.5, ops.PASTE_BOOL)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    by = n * 800 * 1333 + 4 * n * 112 * 112 + 16 * n
    ex['dm_paste_masks'] = {'workload': 'C4 per GPU: 800 instances (8 img x 100 dets), 112x112 -> 800x1333 bool',
                            'ms': ms, 'instances_per_s': n / ms * 1e3, 'algorithmic_bytes': by,
                            'achieved_gbs': by / ms / 1e6, 'frac_of_measured_peak': by / ms / 1e6 / peak}
    # write-only reference point: the same output bytes zero-filled 16 bytes per thread (torch fills a
    # bool tensor one byte per thread, which is far from the write ceiling; an int32 view is not)
    o32 = out.flatten()[:(out.numel() // 4) * 4].view(torch.int32)
    o32.zero_()
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        o32.zero_()
    b.record()
    torch.cuda.synchronize()
    ex['dm_paste_masks']['fill_same_bytes_ms'] = a.elapsed_time(b) / reps
    # inference tail on the same 800 instances, end to end with host results (what mmdet/apis/test.py:54-57
    # needs): (a) reference flow -- paste, copy the N x H x W canvases to the host (RLE would follow on the
    # CPU); (b) fused paste -> RLE on the device, only run boundaries cross PCIe, strings built on the host
    det = torch.cat([boxes, torch.ones(n, 1, device=dev)], 1)
    labels0 = torch.zeros(n, dtype=torch.long, device=dev)

    class _Cfg:
        mask_thr_binary = 0.5
    import time as _time
    for fn, key in ((dm.get_seg_masks, 'get_seg_masks_to_host_ms'), (dm.get_seg_masks_rle, 'get_seg_masks_rle_to_host_ms')):
        # warm-up with two results alive at once, as in the timed loop (`res` is reassigned only after
        # the next call returns), so that the pinned staging blocks both exist before the clock starts
        r1 = fn(logits, det, labels0, _Cfg, (800, 1333, 3), 1.0, False)
        r2 = fn(logits, det, labels0, _Cfg, (800, 1333, 3), 1.0, False)
        del r1, r2
        torch.cuda.synchronize()
        t0 = _time.perf_counter()
        for _ in range(3):
            res = fn(logits, det, labels0, _Cfg, (800, 1333, 3), 1.0, False)
        torch.cuda.synchronize()
        ex['dm_paste_masks'][key] = (_time.perf_counter() - t0) / 3 * 1e3
    ex['dm_paste_masks']['rle_bytes_to_host'] = int(sum(len(r['counts']) for r in res))
    del out, logits, res
    # mask targets: C3 shape, 2 images x 128 positives, all four sizes in one launch
    rng = np.random.default_rng(7 + rank)
    masks_l, props, inds = [], [], []
    for _ in range(2):
        m = synth.make_gt_masks(int(rng.integers(1, 21)), IMG_H, IMG_W, rng)
        pb, pi = synth.jitter_boxes_from_masks(m, 128, rng)
        masks_l.append(dm.BitmapMasks(m, IMG_H, IMG_W))
        props.append(torch.from_numpy(pb).to(dev))
        inds.append(torch.from_numpy(pi).to(dev))
    for _ in range(3):
        t = dm.multi_size_mask_targets(props, inds, masks_l)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        t = dm.multi_size_mask_targets(props, inds, masks_l)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    ex['dm_mask_target'] = {'workload': 'C3: 2 images x 128 positives, G~U{1..20} 800x1344 bitmaps, sizes 14/28/56/112; the '
                                        "batch's bitmaps are packed and uploaded on EVERY call (one pinned staging copy, "
                                        '~10 MB per image) -- the same host arrays each time, so the host side is warm',
                            'ms': ms, 'rois_per_s': 256 / ms * 1e3}
    # the kernel alone on bitmaps that are already resident (single image: the per-object device cache)
    one = dm.BitmapMasks(masks_l[0].masks, IMG_H, IMG_W)
    for _ in range(3):
        t = one.crop_and_resize_device(props[0], [(14, 14), (28, 28), (56, 56), (112, 112)], inds[0], dev, clip=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        t = one.crop_and_resize_device(props[0], [(14, 14), (28, 28), (56, 56), (112, 112)], inds[0], dev, clip=True)
    b.record()
    torch.cuda.synchronize()
    ex['dm_mask_target']['resident_bitmaps_one_image_ms'] = a.elapsed_time(b) / reps
    del t, masks_l

    def timed(fn, reps=10, warm=3):
        for _ in range(warm):
            r = fn()
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            r = fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps, r

    # ---- next rows ---------------------------------------------------------------
    # polygon ground truth (the shipped COCO config): same C3 shape, polygons instead of bitmaps
    polys_l, props, inds = [], [], []
    for _ in range(2):
        objs = synth.make_polygons(int(rng.integers(1, 21)), IMG_H, IMG_W, rng)
        pb, pi = synth.jitter_boxes_from_polygons(objs, 128, rng)
        polys_l.append(dm.PolygonMasks(objs, IMG_H, IMG_W))
        props.append(torch.from_numpy(pb).to(dev))
        inds.append(torch.from_numpy(pi).to(dev))
    ms, _ = timed(lambda: dm.multi_size_mask_targets(props, inds, polys_l))
    ex['dm_polygon_target'] = {'workload': 'C3 with polygon ground truth: 2 images x 128 positives, 1-3 polygons of 5-40 '
                                           'vertices per object, sizes 14/28/56/112, includes the per-step upload',
                               'ms': ms, 'rois_per_s': 256 / ms * 1e3}
    # SimpleRoIAlign at the three SFMStage shapes (100 detections of one 800x1344 image, 256 channels)
    g2 = torch.Generator().manual_seed(5 + rank)
    rois100 = synth.make_rois(1, 100, IMG_H, IMG_W, g2).to(dev)
    sra = {}
    for P, s in ((14, 16), (28, 8), (56, 4)):
        f = torch.randn(1, 256, IMG_H // s, IMG_W // s, device=dev, requires_grad=True)
        # the reference builds every SFMStage with spatial_scale = 1 / semantic_out_stride[-1] = 1/4 while
        # feeding it the stride 16 / 8 / 4 maps (dynamask_head.py:192, :228): measured as it runs there
        layer = dm.SimpleRoIAlign(P, 1.0 / 4)
        ms_f, o = timed(lambda: layer(f, rois100))
        go = torch.ones_like(o)
        ms_b, _ = timed(lambda: ops.simple_roi_align_backward(go, rois100, list(f.shape), 1.0 / 4, True))
        by = o.numel() * 4
        sra['P%d_stride%d' % (P, s)] = {'fwd_ms': ms_f, 'bwd_ms_incl_zero_init': ms_b, 'out_bytes': by,
                                        'fwd_out_gbs': by / ms_f / 1e6}
        del f, o, go
    ex['dm_simple_roi_align'] = {'workload': 'SFMStage gathers: 100 RoIs x 256 ch at 14/28/56 from the stride 16/8/4 maps, '
                                             "every stage with the reference's spatial_scale = 1/4", **sra}
    # fused stage-to-stage refinement, one chunk of 100 detections (28 -> 56 -> 112)
    # every stage predicts the same object: radial blob + N(0,1) noise per pixel at each stage's resolution
    g_st = torch.Generator().manual_seed(77)
    st = [synth.make_mask_logits(100, sz, g_st).to(dev) for sz in (28, 56, 112)]
    ms, _ = timed(lambda: dm.refine_stage_instance_preds(st))
    ex['dm_refine_stages'] = {'workload': '100 detections, stages 28/56/112, in place', 'ms': ms,
                              'instances_per_s': 100 / ms * 1e3}
    # ---- inference tail of one image, device resident, then to host as RLE (config C1/C4 per image):
    # mask RoI extractor 14x14 -> [head convs: PyTorch, not timed] -> refine -> paste -> RLE strings
    feats1 = [torch.randn(1, 256, h, w, device=dev) for (h, w) in synth.pyramid_shapes(IMG_H, IMG_W)]
    ext = dm.SingleRoIExtractor(dict(type='RoIAlign', output_size=14, sampling_ratio=0), 256, STRIDES)
    det100 = torch.cat([rois100[:, 1:], torch.ones(100, 1, device=dev)], 1)
    lab100 = torch.zeros(100, dtype=torch.long, device=dev)

    def tail():
        ins = ext(feats1, rois100)
        final = dm.refine_stage_instance_preds([t.clone() for t in st])
        return ins, dm.get_seg_masks_rle(final, det100, lab100, _Cfg, (800, 1333, 3), 1.0, False)
    tail()
    torch.cuda.synchronize()
    t0 = _time.perf_counter()
    for _ in range(5):
        tail()
    torch.cuda.synchronize()
    dt = (_time.perf_counter() - t0) / 5
    ex['inference_tail_per_image'] = {
        'workload': 'one 800x1333 image, 100 detections: 14x14 mask RoIAlign (256 ch) + refinement 28/56/112 + '
                    'fused paste->RLE, results on the host as RLE strings; head convolutions excluded (PyTorch); stage logits: '
                    'blob + N(0,1) noise at every stage',
        'ms': dt * 1e3, 'img_per_s': 1.0 / dt}
    return ex


# ----------------------------------------------------------------------------------------------
# GPU competitor: the kernels a user of the reference gets on this box today --
# torchvision's CUDA roi_align (the mmcv kernel's lineage, sm_100 SASS shipped in torchvision/_C.so)
# under the reference's per-level select / align / scatter loop
# (mmdet/models/roi_heads/roi_extractors/single_level_roi_extractor.py:53-81).  Library code timed
# beside ours; it is not the oracle and nothing of it is on the product path.
# ----------------------------------------------------------------------------------------------
def tv_single_roi_extractor(feats, rois, out_size, strides, finest_scale=56):
    import torchvision.ops as tvo
    K, C = rois.size(0), feats[0].size(1)
    out = feats[0].new_zeros(K, C, out_size, out_size)
    if len(feats) == 1:
        return tvo.roi_align(feats[0], rois, (out_size, out_size), 1.0 / strides[0], 0, True)
    scale = torch.sqrt((rois[:, 3] - rois[:, 1]) * (rois[:, 4] - rois[:, 2]))
    lvls = torch.floor(torch.log2(scale / finest_scale + 1e-6)).clamp(min=0, max=len(feats) - 1).long()
    for i in range(len(feats)):
        inds = lvls == i
        if inds.any():
            out[inds] = tvo.roi_align(feats[i], rois[inds], (out_size, out_size), 1.0 / strides[i], 0, True)
    return out


def run_competitor(dm, dev, rank, feats, rois, onehot):
    """Ours vs torchvision-CUDA on (a) the C2 step and (b) the extractor calls of one C3 training step
    (2 images: 7x7 bbox features of 1024 RoIs fwd+bwd, 14x14 mask features of 256 positives fwd+bwd,
    56x56 single-level switch input of the same 256 fwd only).  Device-resident, CUDA events."""
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, reps=3, warm=1):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    res = {}
    bucket = onehot.argmax(1)
    fr = [f.detach().requires_grad_() for f in feats]

    def tv_c2():
        for f in fr:
            f.grad = None
        outs = []
        for bi, P in enumerate(BUCKET_SIZES):
            outs.append(tv_single_roi_extractor(fr, rois[bucket == bi], P, STRIDES))
        torch.autograd.backward(outs, [o.detach() for o in outs])
    try:
        res['c2_torchvision_cuda_ms'] = timed(tv_c2, reps=2)
    except Exception as e:  # noqa: BLE001  (e.g. out of memory on a shared box)
        res['c2_torchvision_cuda_error'] = str(e)[:120]
    for f in fr:
        f.grad = None
    torch.cuda.empty_cache()

    # C3: two images of the batch
    f2 = [f[:2].detach().clone().requires_grad_() for f in feats]
    r_bbox = rois[:1024]
    r_mask = torch.cat([rois[:128], rois[512:640]])
    ours7 = dm.SingleRoIExtractor(dict(type='RoIAlign', output_size=7, sampling_ratio=0), 256, STRIDES)
    ours14 = dm.SingleRoIExtractor(dict(type='RoIAlign', output_size=14, sampling_ratio=0), 256, STRIDES)
    ours56 = dm.SingleRoIExtractor(dict(type='RoIAlign', output_size=56, sampling_ratio=0), 256, [4])

    def ours_c3():
        for f in f2:
            f.grad = None
        o7 = ours7(f2, r_bbox)
        o14 = ours14(f2, r_mask)
        o56 = ours56([f2[0].detach()], r_mask)
        torch.autograd.backward([o7, o14], [o7.detach(), o14.detach()])
        return o56

    def tv_c3():
        for f in f2:
            f.grad = None
        o7 = tv_single_roi_extractor(f2, r_bbox, 7, STRIDES)
        o14 = tv_single_roi_extractor(f2, r_mask, 14, STRIDES)
        o56 = tv_single_roi_extractor([f2[0].detach()], r_mask, 56, [4])
        torch.autograd.backward([o7, o14], [o7.detach(), o14.detach()])
        return o56
    res['c3_extractors_ours_ms'] = timed(ours_c3, reps=10, warm=3)
    # the same ~12 launches replayed as one CUDA graph: what the set costs without the host's launch gaps
    try:
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            ours_c3()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            ours_c3()
        res['c3_extractors_ours_cuda_graph_ms'] = timed(graph.replay, reps=10, warm=3)
        del graph
    except Exception as e:  # noqa: BLE001
        res['c3_cuda_graph_error'] = str(e)[:120]
    try:
        res['c3_extractors_torchvision_cuda_ms'] = timed(tv_c3, reps=10, warm=3)
    except Exception as e:  # noqa: BLE001
        res['c3_torchvision_cuda_error'] = str(e)[:120]
    res['workload'] = ('c2: the bench step (8192 RoIs, mixed sizes, fwd+bwd); c3: extractor calls of one training step, '
                       '2 images -- 7x7 x 1024 RoIs fwd+bwd, 14x14 x 256 fwd+bwd, 56x56 single-level x 256 fwd')
    return res


def host_copy_ceiling(world, e2e):
    """The e2e leg is a host <-> device copy problem (37 GB per rank and step against ~15 ms of kernels):
    report the copy rate it reached beside the box's measured ceiling for the same byte mix at the same
    number of ranks (tools/pcie_ceiling.py, table kept in profiles/pcie_ceiling.json)."""
    out = {}
    try:
        per_rank = (e2e['h2d_bytes_per_step'] + e2e['d2h_bytes_per_step']) / (e2e['ms_per_step'] / 1e3) / 1e9
        out['copy_gbs_per_rank'] = per_rank
        with open(os.path.join(ROOT, 'profiles', 'pcie_ceiling.json')) as f:
            table = json.load(f)
        row = table.get(str(world))
        if row:
            out['host_copy_ceiling_gbs_per_rank'] = row['e2e_step_mix_gbs_per_rank']
            out['frac_of_host_copy_ceiling'] = per_rank / row['e2e_step_mix_gbs_per_rank']
            out['host_copy_ceiling_source'] = row.get('source', 'profiles/pcie_ceiling.json')
    except Exception:
        pass
    return out


def run_e2e(args, dm, dev, rank, world, feats_dev, rois_h, onehot_h, counts):
    """Same pass through the plugin surface (BucketedRoIExtractor + autograd) with HOST buffers.

    The batch is fed in chunks of `--e2e-chunk-images` images so that the pinned staging buffers
    stay a few GB per rank (8 ranks share one host); every byte of the pyramid, the pooled
    features and the gradient pyramid still crosses PCIe inside the timed region."""
    import torch.distributed as dist
    C = args.channels
    ci = max(1, min(args.e2e_chunk_images, args.batch))
    R = args.rois_per_img
    ext = dm.BucketedRoIExtractor(dict(type='RoIAlign', output_size=14, sampling_ratio=0), C, STRIDES)
    bucket_h = onehot_h.argmax(1)
    chunks = []
    for i0 in range(0, args.batch, ci):
        i1 = min(i0 + ci, args.batch)
        sel = slice(i0 * R, i1 * R)
        r = rois_h[sel].clone()
        r[:, 0] -= i0
        cnt = torch.bincount(bucket_h[sel], minlength=len(BUCKET_SIZES)).tolist()
        chunks.append((i0, i1, r.pin_memory(), onehot_h[sel].clone().pin_memory(), cnt))
    max_cnt = [max(c[4][b] for c in chunks) for b in range(len(BUCKET_SIZES))]
    NBUF = 2   # result staging is double buffered: chunk i's D2H overlaps chunk i+1's H2D and kernels
    try:
        feats_pin = [torch.empty(f.shape, dtype=f.dtype, pin_memory=True) for f in feats_dev]
        for p, f in zip(feats_pin, feats_dev):
            p.copy_(f)
        outs_pin = [[torch.empty((max_cnt[b], C, s, s), dtype=torch.float32, pin_memory=True)
                     for b, s in enumerate(BUCKET_SIZES)] for _ in range(NBUF)]
        grads_pin = [[torch.empty((ci, ) + tuple(f.shape[1:]), dtype=f.dtype, pin_memory=True) for f in feats_dev]
                     for _ in range(NBUF)]
    except RuntimeError as e:
        return {'value': None, 'unit': UNIT, 'error': 'pinned allocation failed: %s' % str(e)[:80]}
    h2d = sum(p.numel() * 4 for p in feats_pin) + rois_h.numel() * 4 + onehot_h.numel() * 4
    d2h = sum(counts[b] * C * s * s * 4 for b, s in enumerate(BUCKET_SIZES)) + sum(p.numel() * 4 for p in feats_pin)
    main = torch.cuda.current_stream(dev)
    copy = torch.cuda.Stream(dev)    # device -> host results; PCIe is full duplex, so it runs beside the H2D
    torch.cuda.synchronize()

    def one():
        done = [None] * NBUF          # event: the host may read / reuse staging set k
        for j, (i0, i1, r_pin, o_pin, cnt) in enumerate(chunks):
            k = j % NBUF
            if done[k] is not None:
                done[k].synchronize()     # the host consumes that set's results before it is refilled
            fd = [p[i0:i1].to(dev, non_blocking=True).requires_grad_() for p in feats_pin]
            rd = r_pin.to(dev, non_blocking=True)
            od = o_pin.to(dev, non_blocking=True)
            res = ext.forward_bucketed(fd, rd, od)
            outs = [o.detach() for o in res.feats]
            fwd_done = torch.cuda.Event()
            fwd_done.record(main)
            with torch.cuda.stream(copy):
                copy.wait_event(fwd_done)
                for p, o in zip(outs_pin[k], outs):
                    p[:o.size(0)].copy_(o, non_blocking=True)
                    o.record_stream(copy)
            torch.autograd.backward(res.feats, outs)
            bwd_done = torch.cuda.Event()
            bwd_done.record(main)
            with torch.cuda.stream(copy):
                copy.wait_event(bwd_done)
                for p, f in zip(grads_pin[k], fd):
                    p[:i1 - i0].copy_(f.grad, non_blocking=True)
                    f.grad.record_stream(copy)
                done[k] = torch.cuda.Event()
                done[k].record(copy)
        torch.cuda.synchronize()

    one()  # warm-up
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        one()
    if world > 1:
        dist.barrier()
    dt = (time.perf_counter() - t0) / args.e2e_steps
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    return {'value': world * rois_h.size(0) / dt, 'unit': UNIT, 'ms_per_step': dt * 1e3,
            'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h), 'steps': args.e2e_steps,
            'chunk_images': ci,
            'api': 'BucketedRoIExtractor.forward_bucketed + autograd backward per chunk, pinned host in / out, '
                   'results double buffered on a copy stream'}


if __name__ == '__main__':
    main()
