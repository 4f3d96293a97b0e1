#!/usr/bin/env python
"""bench.py -- V2CE hot path on B200: 346x260 frame-pairs/s (and LDATI Mevents/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): center 346x260, 321 synthetic gray frames = 20 windows of
16 frame pairs, batch 4, random-init V2ce3d, voxel + LDATI + event-frame path.  One *step* is one
batch of 4 windows (64 frame pairs) through v2ce_toolbox_b200.runner.BatchRunner: UNet ->
event-frame accumulate/select/normalise -> LDATI count/emit/sort/pack, the last two stages on a second
stream under the UNet of the next step.  Step i uses batch (i mod 5) of the clip; at N GPUs every rank
runs its own windows (weak scaling, windows are independent) and the per-rank event shards are gathered
to rank 0 over NCCL inside the timed region.

  value  device-timed throughput, inputs (float32 image units) already resident in HBM, results left on
         the device
  e2e    same steps with HOST buffers: raw uint8 frame windows in pinned memory H2D (the pre-processing
         runs inside the head conv), packed events + preview frames D2H, inside the timed region
  roofline    tensor roofline of the UNet forward (2169.336 GFLOP per window) against
              the measured sustained bf16 peak: CUDA events around the forward run alone over the same
              K steps (`forward_ms`) and inside the timed steps (`forward_ms_in_step`)
  ldati  LDATI-only microbench on a bounded sample of configs[2] (24 pairs, two voxel distributions)
  cpu_baseline  the CPU oracle (torch-CPU fp32 UNet + numpy LDATI/EF restatement, oracle/) timed on
              the host cores on a bounded sample (1 window = 16 frame pairs)

--impl reference times that CPU oracle port alone (the reference itself is Python and is not
present on the GPU box; oracle/ is pinned to it by tests/golden).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, L = 260, 346, 16
N_FRAMES = 321
BATCH = 4
GFLOP_PER_WINDOW = 2169.336
METRIC = '346x260 frame-pairs/s (center, 321 frames, batch 4, voxel+LDATI+event-frame)'
# dram__bytes_read.sum + dram__bytes_write.sum of one V2ce3d forward (batch 4), summed over its launches from the
# ncu capture summarised in profiles/forward_traffic_r1.txt
TRAFFIC_BYTES = 13.03e9        # profiles/forward_traffic_r1_b.txt: read 8.98 GB + write 4.05 GB (L2: 59.1 GB)
KERNEL_NOTE = ('V2ce3d forward (26 launches + 4 spectral-norm launches on a side stream): conv_halo_kdm_kernel x11 '
               '(head, stride-2 encoder convs and decoder convs with fused shortcuts, N<=64 convs), conv_halo_kernel x10, '
               'conv_igemm_kernel x4 (remaining 1x1x1 shortcuts, side stream), head prep')


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get('bf16_tflops_sustained', 1390.9), d.get('hbm_gbs', 6531.6), 'measured'
    return 1400.0, 6650.0, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        busy = [s for s in sm if s > 0]
        return {'sm_mhz': float(np.median(busy)) if busy else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def make_inputs():
    """The 20 windows of the clip: preprocessed image units (20, 16, 2, 260, 346) float32 and the raw gray frames
    they come from (20, 17, 260, 346) uint8 (both host)."""
    import synth_inputs as synth
    from v2ce_toolbox_b200.v2ce import image_pre_processing, window_schedule
    frames = synth.make_video(N_FRAMES, H, W, seed=0)
    starts, mode = window_schedule(N_FRAMES, L)
    assert mode == 0 and len(starts) == 20
    units = torch.stack([image_pre_processing(frames[s:s + L + 1], H) for s in starts], dim=0)
    windows = torch.from_numpy(np.stack([frames[s:s + L + 1] for s in starts], axis=0))     # (20, 17, H, W) uint8
    return units, windows


# ------------------------------------------------------------------------------------------------
# CPU oracle arm (cpu_baseline and --impl reference)
# ------------------------------------------------------------------------------------------------
def cpu_port_step(orc, units_window, pair_base, fps=30):
    from oracle import ef_oracle, ldati_oracle
    vox = orc.forward(units_window).numpy().reshape(-1, 2, 10, H, W)
    frames, _, _ = ef_oracle.event_frames_oracle(vox, 10, 98, True)
    ev = ldati_oracle.sample_voxel_statistical_oracle(vox, fps=fps, seed=0, frame_base=pair_base, flavor='cpu')
    return sum(len(e) for e in ev)


def time_cpu_port(units, steps, warmup):
    import synth_inputs as synth
    from oracle.unet_oracle import UNetOracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    orc = UNetOracle(synth.make_state_dict(0, 'reference'))
    for i in range(warmup):
        cpu_port_step(orc, units[i % 20:i % 20 + 1], 0)
    t0 = time.perf_counter()
    events = 0
    for i in range(steps):
        events += cpu_port_step(orc, units[i % 20:i % 20 + 1], (i % 20) * L)
    dt = time.perf_counter() - t0
    return steps * L / dt, events / dt / 1e6, dt, cores


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    units, _ = make_inputs()
    # bounded: one window (16 pairs) per step; cap the step count so the arm ends within minutes
    steps = max(1, min(args.steps, 6))
    warmup = max(1, min(args.warmup, 1))
    pairs_s, mev_s, dt, cores = time_cpu_port(units, steps, warmup)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': pairs_s, 'unit': 'frame-pairs/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warmup, 'ms_per_step': dt / steps * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'center 346x260, 321 frames, batch 4, voxel+LDATI+event-frame video',
                   'sample': '1 window (16 frame pairs) per step on host cores'},
        'cpu_baseline': {'value': pairs_s, 'unit': 'frame-pairs/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{steps} steps x 1 window (16 pairs): torch-CPU fp32 UNet + numpy LDATI/EF oracle'},
        'e2e': {'value': pairs_s, 'unit': 'frame-pairs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'ldati_mevents_per_s': mev_s, 'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
class Runner:
    """Model + two BatchRunners (v2ce_toolbox_b200.runner): device-resident inputs / results for `value`,
    pinned host inputs and host results for `e2e`."""

    def __init__(self, device, units_host, windows_host, rank, world):
        import synth_inputs as synth
        from v2ce_toolbox_b200.runner import BatchRunner
        from v2ce_toolbox_b200.scripts.v2ce_3d import V2ce3d
        self.device = device
        self.rank, self.world = rank, world
        self.model = V2ce3d()
        self.model.load_state_dict(synth.make_state_dict(0, 'reference'))
        self.model.eval().to(device)
        self.dev_runner = BatchRunner(self.model, device, fps=30, seed=0, copy_out=False)
        self.host_runner = BatchRunner(self.model, device, fps=30, seed=0, copy_out=True)
        self.dev_runner.time_forward = True
        # value: image units resident in HBM; e2e: the raw uint8 windows in pinned host memory (the pre-processing
        # of v2ce.py:45-64 runs inside the head conv, V2ce3d.forward_frames)
        self.units_dev = [units_host[i:i + BATCH].contiguous().to(device) for i in range(0, 20, BATCH)]
        self.units_pinned = [windows_host[i:i + BATCH].contiguous().pin_memory() for i in range(0, 20, BATCH)]
        self.n_pairs = BATCH * L
        self.comm_stream = None

    def gather(self, br, t):
        """Final gather of the per-rank event shards to rank 0 (NCCL over NVLink), on the post stream, behind
        the pack kernel of batch t."""
        import torch.distributed as dist
        if self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream(device=self.device)
        with torch.cuda.stream(self.comm_stream):
            # its own stream: the post stream already holds stage A of the next batch, which waits for the next UNet
            self.comm_stream.wait_event(t.packed)
            total = t.total
            cnt = torch.tensor([total], dtype=torch.int64, device=self.device)
            counts = [torch.zeros(1, dtype=torch.int64, device=self.device) for _ in range(self.world)]
            dist.all_gather(counts, cnt)
            mx = int(max(int(c.item()) for c in counts))
            pad = torch.zeros(mx * 13, dtype=torch.uint8, device=self.device)
            pad[:total * 13] = t.events_dev[:total * 13]
            if self.rank == 0:
                bufs = [torch.empty(mx * 13, dtype=torch.uint8, device=self.device) for _ in range(self.world)]
                dist.gather(pad, bufs, dst=0)
            else:
                dist.gather(pad, None, dst=0)
            fin = torch.cuda.Event()
            fin.record(self.comm_stream)
        return fin


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    device = torch.device('cuda', local_rank)
    torch.cuda.set_device(device)
    affinity = None
    if world > 1:
        # one process per GPU: keep each rank (and the pinned staging it allocates) on its GPU's NUMA node
        from v2ce_toolbox_b200.dist import bind_to_gpu_numa
        affinity = bind_to_gpu_numa(local_rank)
    units, windows = make_inputs()
    r = Runner(device, units, windows, rank, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(host_io):
        """One step = one batch of 4 windows through BatchRunner.submit; the results of step i are collected
        (and, at N > 1, gathered to rank 0) while step i+1 computes."""
        br = r.host_runner if host_io else r.dev_runner
        src = r.units_pinned if host_io else r.units_dev
        stats = {'events': 0, 'fwd': [], 'h2d': 0, 'd2h': 0}

        def collect(t, timed):
            br.wait(t, copy=False)
            if world > 1:
                r.gather(br, t).synchronize()
            if timed:
                stats['events'] += t.total
                stats['h2d'], stats['d2h'] = t.h2d_bytes, t.d2h_bytes
                if t.fwd_events is not None:
                    stats['fwd'].append(t.fwd_events[0].elapsed_time(t.fwd_events[1]))

        def run(n, timed):
            prev = None
            for i in range(n):
                b = i % 5
                t = br.submit(src[b], (rank * 5 + b) * r.n_pairs)
                if prev is not None:
                    collect(prev, timed)
                prev = t
            if prev is not None:
                collect(prev, timed)

        run(args.warmup, False)
        br.launches = 0
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        s.record()
        run(args.steps, True)
        torch.cuda.synchronize()
        e.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = max(s.elapsed_time(e), wall * 1e3) if host_io else s.elapsed_time(e)
        tt = torch.tensor([ms], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item()), stats, br.launches

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, st_dev, launches = timed_loop(False)
    ms_e2e, st_e2e, _ = timed_loop(True)
    # the network alone (no event-frame / LDATI kernels sharing the SMs): what the roofline fraction describes
    fwd_ms = []
    for i in range(args.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r.model(r.units_dev[i % 5])
        e1.record()
        fwd_ms.append((e0, e1))
    torch.cuda.synchronize()
    fwd_alone = float(np.mean([a.elapsed_time(b) for a, b in fwd_ms]))
    clocks = sampler.stop() if rank == 0 else None

    pairs_per_step = r.n_pairs * world
    value = pairs_per_step * args.steps / (ms_dev / 1e3)
    e2e = pairs_per_step * args.steps / (ms_e2e / 1e3)
    if rank != 0:
        return
    tf_peak, hbm_peak, how = peaks()
    fwd_in_step = float(np.mean(st_dev['fwd'])) if st_dev['fwd'] else float('nan')
    fwd = fwd_alone
    achieved = GFLOP_PER_WINDOW * BATCH / fwd            # GFLOP / ms = TFLOP/s
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        pairs_s, mev_s, dt, cores = time_cpu_port(units, 2, 1)
        cpu = {'value': pairs_s, 'unit': 'frame-pairs/s', 'cores': cores, 'kind': 'port',
               'sample': '2 steps x 1 window (16 pairs): torch-CPU fp32 UNet + numpy LDATI/EF oracle',
               'ldati_mevents_per_s': mev_s}
    line = {
        'metric': METRIC, 'value': value, 'unit': 'frame-pairs/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_dev / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
        'config': {'workload': 'center 346x260, 321 frames, batch 4, voxel+LDATI+event-frame video',
                   'pairs_per_step_per_gpu': r.n_pairs, 'weights': 'random-init V2ce3d (seed 0)',
                   'l2': 'inputs+activations per step (2.6 GB) exceed the 126 MB L2',
                   'pipeline': 'event frames + LDATI of step i run on a second stream under the UNet of step i+1',
                   'e2e_input': 'uint8 gray frame windows (4 x 17 x 260 x 346) in pinned host memory; pre-processing '
                                'fused into the head conv; events + preview frames copied back to pinned host memory',
                   'multi_gpu': 'windows sharded per rank, NCCL gather of event shards to rank 0'},
        'e2e': {'value': e2e, 'unit': 'frame-pairs/s', 'h2d_bytes_per_step': st_e2e['h2d'],
                'd2h_bytes_per_step': st_e2e['d2h'], 'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': launches,
        'ldati_mevents_per_s': st_dev['events'] * world / (ms_dev / 1e3) / 1e6,
        'events_per_pair': st_dev['events'] / (r.n_pairs * args.steps),
        'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': tf_peak, 'unit': 'TFLOP/s',
                     'frac': achieved / tf_peak, 'traffic': TRAFFIC_BYTES, 'peak_source': how,
                     'kernel': KERNEL_NOTE,
                     'forward_ms': fwd, 'forward_ms_in_step': fwd_in_step,
                     'frac_in_step': GFLOP_PER_WINDOW * BATCH / fwd_in_step / tf_peak,
                     'forward_share_of_step': fwd / (ms_dev / args.steps),
                     'note': 'forward_ms: CUDA events around the network run alone over the same K steps; '
                             'forward_ms_in_step: the same events inside the timed steps, where the previous '
                             'batch\'s event-frame / LDATI kernels share the SMs'},
        'cpu_baseline': cpu,
        'clocks': clocks,
        'cpu_affinity': None if affinity is None else {'cpus': len(affinity), 'first': affinity[0], 'last': affinity[-1]},
    }
    if world == 1:
        try:
            micro = ldati_microbench(device, hbm_peak)
        except Exception as e:                      # the secondary table must never cost the headline line
            micro = {'error': f'{type(e).__name__}: {e}'}
        line['ldati'] = {'workload': 'LDATI-only microbench, 346x260x10 bins, 24 and 96 pairs per call (bounded sample of '
                                     'BASELINE configs[2]); includes the host read of the counts between the two phases',
                         'hbm_peak_gbs': hbm_peak, **micro}
    print(json.dumps(line), flush=True)


def ldati_microbench(device, hbm_peak, reps=5, pairs=(24, 96)):
    """BASELINE.json configs[2] on a bounded sample: LDATI alone on synthetic event-count voxels, distributions
    (a) torch.rand and (b) randint(0,10) taken from the reference's own bench (LDATI.py:327-346) and (c) 0.015*rand
    (sparse), in calls of 24 frame
    pairs (the reference's stage-2 chunk, v2ce.py:301 -- its dense (B,2,9,H,W,M) tensors do not fit more) and of 96
    pairs (this path holds one 4-byte word per event, so the chunk is only bounded by 2^31 events per call).
    Device-timed count -> (counts D2H) -> emit/sort/pack; HBM roofline with SURVEY.md 8d's algorithmic bytes:
    4*20*H*W per pair read + 13 B per event written."""
    from v2ce_toolbox_b200 import ldati
    eng = ldati.LdatiEngine(device)
    out = {}
    for F in pairs:
        # (c) 0.015 * rand: the sparse, all-ties regime of a random-init network, 24 pairs only
        for name in ('rand', 'randint10') + (('sparse',) if F == 24 else ()):
            g = torch.Generator(device=device).manual_seed(42)
            if name == 'rand':
                vox = torch.rand((F, 2, 10, H, W), generator=g, device=device)
            elif name == 'sparse':
                vox = torch.rand((F, 2, 10, H, W), generator=g, device=device) * 0.015
            else:
                vox = torch.randint(0, 10, (F, 2, 10, H, W), generator=g, device=device).float()
            params = ldati.make_params(F, H, W, fps=30, seed=42, frame_base=0, device=device)
            total = 0
            for _ in range(2):
                ev, seg, st = eng.run(vox, params)
                total = int(seg.sum())
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                eng.run(vox, params)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            alg = F * 4 * 20 * H * W + 13 * total
            out[name if F == 24 else f'{name}_{F}pairs'] = {
                'pairs': F, 'events_per_pair': total / F, 'ms': ms, 'mevents_per_s': total / ms / 1e3,
                'pairs_per_s': F / ms * 1e3, 'algorithmic_bytes': alg, 'achieved_gbs': alg / ms / 1e6,
                'frac_of_hbm_peak': alg / ms / 1e6 / hbm_peak}
            del ev, vox
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference_arm(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the V2CE B200 path has no CPU fallback '
                         '(use --impl reference for the CPU oracle arm)')
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
