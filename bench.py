#!/usr/bin/env python
"""bench.py -- images/sec of Darknet19-YOLO2 416x416 inference (forward + region decode + per-class
NMS), the metric BASELINE.json names, on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W]              # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]        # CPU restatement of the reference
    torchrun ... bench.py --gpus N ...                               # N > 1: one rank per GPU

One "step" = one pass of the hot path over one synthetic batch (64 images of 416x416x3 uint8 per
GPU, random-init weights of the reference's initialiser, core BN in inference mode, head BN in
batch-statistics mode -- the reference detect script's graph).  Prints ONE JSON line (rank 0).

value      device-timed throughput with the batch already resident in HBM (CUDA events per step on
           the launching stream, L2 flushed between steps, max over ranks).
e2e        same metric through Yolo2Engine.submit(): pinned host uint8 batch -> H2D -> step -> D2H of
           the detections, every copy inside the timed region (the H2D of batch i+1 runs on a copy
           stream and overlaps the kernels of batch i, as a serving loop would).
roofline   tensor-core bound: algorithmic conv FLOPs of one step / conv time per step vs MEASURED_PEAKS.json.
           Conv time = min(sum of CUDA events around each of the 22 conv launches in an eager replay, the whole
           device-timed graph step): eager launches leave idle gaps that the events count, and the convs cannot take
           longer than the step that contains them (roofline.conv_ms_source names the bound used).  roofline.traffic =
           DRAM bytes of those launches from the committed ncu launch list (profiles/*_traffic.json).
cpu_baseline  the oracle (CPU restatement of the reference, PyTorch-CPU fp32) on a bounded sample.

    --image-size 608 --batch 32    BASELINE.json configs[3] (19x19 grid) instead of the headline 416 / 64
    --no-graph                     eager launches (for ncu launch lists);  --no-cpu-baseline skips the CPU leg
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH_PER_GPU = 64
IMAGE_SIZE = 416
OUTPUT_FILTER = 125
SCORE_THRESH, IOU_THRESH = 0.3, 0.45
METRIC = 'images/sec Darknet19-YOLO2 416x416 (fwd+decode+NMS)'

# (k, cin, cout, out_hw) at 416^2 -> algorithmic FLOPs = 2*k*k*cin*cout*hw*hw   
def conv_flops_per_image(image_size, output_filter):
    from tensorflow_yolo2_b200.yolo2_nets.darknet import CORE_PLAN
    h = image_size
    total, per = 0.0, []
    plan = list(CORE_PLAN) + [(3, 1024, 1024, False)] * 3 + [(1, 1024, output_filter, False)]
    for (k, cin, cout, pool) in plan:
        f = 2.0 * k * k * cin * cout * h * h
        per.append(f)
        total += f
        if pool:
            h //= 2
    return total, per


def make_config(world, extra=None):
    """config of the bench line -- the reference arm prints the same one (it is the driver's join key)."""
    cfg = dict(workload='Darknet19-YOLO2 %dx%d inference, fwd + region decode + per-class NMS, synthetic uint8 batch ' % (IMAGE_SIZE, IMAGE_SIZE) +
                        '%d per GPU (BASELINE.json configs[%d])' % (BATCH_PER_GPU, 1 if IMAGE_SIZE == 416 else 3),
               global_batch=world * BATCH_PER_GPU, image_size=IMAGE_SIZE, output_filter=OUTPUT_FILTER,
               score_thresh=SCORE_THRESH, iou_thresh=IOU_THRESH, head_bn='batch statistics',
               l2='flushed (256 MiB memset) between timed steps', parallelism='batch sharding x%d' % world)
    if extra:
        cfg.update(extra)
    return cfg


def load_traffic():
    """DRAM bytes (read + write) of the 22 conv launches of one step, from the committed ncu launch list of this same
    command (profiles/r1f_traffic.json; ncu numbers are never taken live inside a timed run)."""
    p = os.path.join(ROOT, 'profiles', 'r1f_traffic.json')
    try:
        return float(json.load(open(p))['conv_dram_bytes_per_step'])
    except Exception:  # noqa: BLE001
        return None


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(burst=d.get('bf16_tflops', 1590.0), sustained=d.get('bf16_tflops_sustained', 1400.0),
                    hbm=d.get('hbm_gbs', 6650.0), which='measured')
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, which='fallback')


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region, sampled in-process through NVML every 20 ms (the
    B200_PROFILING.md recipe's nvidia-smi query, without a subprocess' start-up latency)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.maxclk = index, [], set(), None
        self._stop = None
        self._thr = None

    def start(self):
        import threading
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                import torch
                uuid = 'GPU-' + str(torch.cuda.get_device_properties(self.index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.maxclk = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            return
        names = dict(hw_slowdown=0x8, hw_thermal_slowdown=0x40, sw_thermal_slowdown=0x20, sw_power_cap=0x4)   # NVML reason bits
        self._stop = threading.Event()

        def loop():
            while not self._stop.is_set():
                try:
                    self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                    try:
                        r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                    except Exception:
                        r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for k, bit in names.items():
                        if r & bit:
                            self.reasons.add(k)
                except Exception:
                    pass
                self._stop.wait(0.02)
        self._thr = threading.Thread(target=loop, daemon=True)
        self._thr.start()

    def stop(self):
        if self._thr is None:
            return None
        self._stop.set()
        self._thr.join(timeout=2)
        if not self.samples:
            return None
        return dict(sm_mhz=statistics.median(self.samples), sm_min_mhz=min(self.samples), sm_max_mhz=self.maxclk,
                    reasons=sorted(self.reasons), samples=len(self.samples))


# ------------------------------------------------------------------------------------------------
# CPU restatement of the reference (oracle) -- used ONLY for cpu_baseline / --impl reference
# ------------------------------------------------------------------------------------------------
def cpu_reference_step(batch, threads=None, reps=1, seed=0):
    import numpy as np
    import torch
    from oracle import yolo2_oracle as O
    from tests.helpers import make_store, oracle_params
    if threads:
        torch.set_num_threads(threads)
    st, layers = make_store(OUTPUT_FILTER, seed=seed)
    core_p, head_p = oracle_params(st, layers)
    img = np.random.RandomState(seed).randint(0, 256, (batch, IMAGE_SIZE, IMAGE_SIZE, 3)).astype(np.uint8)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        x = torch.tensor(O.preprocess_u8(img))
        with torch.no_grad():
            net = O.darknet19_forward(x, core_p, head_p, core_training=False, head_training=True,
                                      dtype=torch.float32).numpy()
        boxes, sthr, _ = O.region_decode_v2(net, O.VOC_ANCHORS, 20, SCORE_THRESH)
        for n in range(batch):
            O.nms_per_class(boxes[n], sthr[n], IOU_THRESH, SCORE_THRESH)
        times.append(time.perf_counter() - t0)
    return times, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample_batch = 8
    t_all, threads = cpu_reference_step(sample_batch, threads=cores, reps=args.warmup + args.steps)
    timed = t_all[args.warmup:]
    total = sum(timed)
    value = sample_batch * len(timed) / total
    sample = 'batch %d of %dx%d per step (bounded sample of the batch-64 workload), torch-CPU fp32' % (sample_batch, IMAGE_SIZE, IMAGE_SIZE)
    line = dict(metric=METRIC.replace('416x416', '%dx%d' % (IMAGE_SIZE, IMAGE_SIZE)), value=value, unit='images/s', n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * total / len(timed), higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='f32', data='synthetic', impl='reference',
                config=make_config(max(args.gpus, 1)),
                cpu_baseline=dict(value=value, unit='images/s', cores=threads, kind='port', sample=sample),
                e2e=dict(value=value, unit='images/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from tensorflow_yolo2_b200 import ops
    from tensorflow_yolo2_b200.engine import Yolo2Engine

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    N = BATCH_PER_GPU
    eng = Yolo2Engine(N, IMAGE_SIZE, OUTPUT_FILTER, score_thresh=SCORE_THRESH, iou_thresh=IOU_THRESH, max_keep=64,
                      use_cuda_graph=not args.no_graph, device=dev, seed=0)
    # 4 distinct synthetic batches (4 x 33 MB > L2 is not needed: L2 is flushed between steps anyway)
    g = torch.Generator(device='cpu').manual_seed(1234 + rank)
    host_batches = [torch.randint(0, 256, (N, IMAGE_SIZE, IMAGE_SIZE, 3), dtype=torch.uint8, generator=g).pin_memory()
                    for _ in range(2)]
    dev_batches = [b.to(dev) for b in host_batches]
    flush = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)      # > 126 MB L2
    stream = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(i, timed):
        eng.in_u8.copy_(dev_batches[i % len(dev_batches)])       # device->device, outside the events
        flush.zero_()                                            # L2 flush, outside the events
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            eng.run()
            e1.record(stream)
            return e0, e1
        eng.run()
        return None

    for i in range(max(args.warmup, 3)):
        one_step(i, False)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n_launch0 = ops.launch_count()
    evs = [one_step(i, True) for i in range(args.steps)]
    eager_launches = ops.launch_count() - n_launch0          # 0 under CUDA-graph replay (counted at capture)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t_ms = sum(a.elapsed_time(b) for a, b in evs)
    tt = torch.tensor([t_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_ms = float(tt.item())
    value = world * N * args.steps / (t_ms * 1e-3)
    cand = int((eng.scores > 0).sum().item())
    kept = int(eng.keep_count.sum().item())

    # ---- e2e: pinned host uint8 -> H2D -> step -> D2H (keep lists + boxes of the batch) ----
    res_host = dict(keep_idx=torch.empty(eng.keep_idx.shape, dtype=torch.int32).pin_memory(),
                    keep_count=torch.empty(eng.keep_count.shape, dtype=torch.int32).pin_memory(),
                    boxes=torch.empty(eng.boxes.shape, dtype=torch.float32).pin_memory())
    h2d = host_batches[0].numel()
    d2h = sum(t.numel() * t.element_size() for t in res_host.values())

    def e2e_step(i):
        eng.submit(host_batches[i % len(host_batches)], res_host)     # H2D (copy stream) | step | D2H, pipelined

    for i in range(3):
        e2e_step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        e2e_step(i)
    e1.record(stream)
    barrier()
    te = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * N * args.steps / (float(te.item()) * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (conv_tc_kernel), timed live per launch ----
    launches_per_step = eng.launches_per_step if not args.no_graph else eager_launches // max(args.steps, 1)
    eng.use_cuda_graph = False
    conv_ms = conv_kernel_times(eng, ops, iters=max(3, min(args.steps, 10)))
    flops_img, per_layer = conv_flops_per_image(IMAGE_SIZE, OUTPUT_FILTER)
    peaks = load_peaks()
    conv_total_ms = sum(conv_ms)
    conv_ms_source = 'CUDA events around each of the 22 conv launches (eager replay of the step)'
    ms_per_step = t_ms / args.steps
    if not args.no_graph and conv_total_ms > ms_per_step:
        # eager launches leave the GPU idle between kernels and the events count that gap; the convs cannot take longer
        # than the whole graph-replayed step that contains them (device-timed above), so that is the tighter bound
        conv_total_ms = ms_per_step
        conv_ms_source = 'whole device-timed step (upper bound: the per-launch events of the eager replay summed to more)'
    achieved = N * flops_img / (conv_total_ms * 1e-3) / 1e12
    roofline = dict(bound='tensor', achieved=achieved, peak=peaks['sustained'], unit='TFLOP/s',
                    frac=achieved / peaks['sustained'], traffic=load_traffic() if IMAGE_SIZE == 416 and N == 64 else None,
                    algorithmic_bytes=N * (35.0e6 if IMAGE_SIZE == 416 else 35.0e6 * IMAGE_SIZE * IMAGE_SIZE / (416.0 * 416.0)), traffic_source='profiles/r1f_traffic.json (ncu launch list of this command)', peak_source=peaks['which'] + ' sustained bf16',
                    kernel='conv_tc_kernel x21 + conv1_u8_pool_kernel (22 conv launches/step)', conv_ms_per_step=conv_total_ms, conv_ms_source=conv_ms_source,
                    per_layer_tflops=[round(N * f / (ms * 1e-3) / 1e12, 1) for f, ms in zip(per_layer, conv_ms)])

    # ---- CPU baseline: oracle on a bounded sample ----
    cpu = None
    if not args.no_cpu_baseline:
        try:
            times, threads = cpu_reference_step(4, threads=os.cpu_count(), reps=2)
            cpu = dict(value=4 / min(times), unit='images/s', cores=threads, kind='port',
                       sample='batch 4 of 416x416 (fwd+decode+NMS), best of 2, torch-CPU fp32 restatement of the reference')
        except Exception as e:  # noqa: BLE001
            cpu = dict(value=None, unit='images/s', cores=os.cpu_count(), kind='port', sample='failed: %r' % (e,))

    line = dict(metric=METRIC.replace('416x416', '%dx%d' % (IMAGE_SIZE, IMAGE_SIZE)), value=value, unit='images/s', n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                ms_per_step=t_ms / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='bf16',
                data='synthetic',
                config=make_config(world, dict(nms_candidates=cand, nms_kept=kept)),
                e2e=dict(value=e2e_value, unit='images/s', h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
                gpu_launches=int(launches_per_step * args.steps), launches_per_step=int(launches_per_step),
                roofline=roofline, cpu_baseline=cpu, clocks=clocks)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def conv_kernel_times(eng, ops, iters=5):
    """Per-layer conv kernel durations (ms) with CUDA events around each conv launch, eager mode."""
    import torch
    real, real1 = ops.conv_fwd_bf16, ops.conv1_u8_pool
    records = []

    def timed(fn):
        def wrapper(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            records.append((e0, e1))
            return out
        return wrapper

    ops.conv_fwd_bf16, ops.conv1_u8_pool = timed(real), timed(real1)     # the engine calls them in layer order
    try:
        eng._enqueue()
        torch.cuda.synchronize()
        records.clear()
        for _ in range(iters):
            eng._enqueue()
        torch.cuda.synchronize()
    finally:
        ops.conv_fwd_bf16, ops.conv1_u8_pool = real, real1
    nl = len(eng.layers)
    ms = [0.0] * nl
    for i, (a, b) in enumerate(records):
        ms[i % nl] += a.elapsed_time(b) / iters
    return ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='eager launches (for ncu launch lists)')
    ap.add_argument('--image-size', type=int, default=None, help='416 (headline, BASELINE configs[1]) or 608 (configs[3])')
    ap.add_argument('--batch', type=int, default=None, help='images per GPU (default 64; configs[3] uses 32 at 608)')
    args = ap.parse_args()
    global IMAGE_SIZE, BATCH_PER_GPU
    if args.image_size:
        IMAGE_SIZE = args.image_size
    if args.batch:
        BATCH_PER_GPU = args.batch
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
