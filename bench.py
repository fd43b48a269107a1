#!/usr/bin/env python
"""bench.py — learner samples/sec of the AlphaStar policy hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]                  # our CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]  # the CPU oracle port of the reference
    torchrun ... bench.py --gpus N ...                                    # one rank per GPU (driver launches this)

A step = ``RLLearner._train`` on one synthetic learner batch: rl_learner_forward -> ReinforcementLoss ->
backward -> (N>1: one NCCL all-reduce of the flat gradient arena) -> clip -> Adam.  Workload at N=1 is
BASELINE.json configs[3]/metric: per-rank batch 128, unroll 32, 512 entities, 128x128 spatial (weak scaling:
every rank trains its own 128 trajectories).  `value` = world * B * T / step_time with the batch resident in HBM;
`e2e` = same through the public API with the batch copied from pinned host memory every step and the loss read
back.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=128, help='trajectories per rank (BASELINE: 128)')
    ap.add_argument('--unroll', type=int, default=32, help='unroll length (BASELINE: 32)')
    ap.add_argument('--encoder-chunk', type=int, default=264)
    ap.add_argument('--terms', type=int, default=3, help='tensor-core products per GEMM: 3 = fp32-class (parity), 1 = bf16')
    ap.add_argument('--no-checkpoint', action='store_true', help='keep encoder activations instead of recomputing them')
    ap.add_argument('--keep-chunks', type=int, default=16,
                    help='number of encoder chunks whose entity-transformer activations are kept (not recomputed)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--cpu-batch', type=int, default=4)
    ap.add_argument('--cpu-unroll', type=int, default=8)
    ap.add_argument('--cpu-threads', type=int, default=0,
                    help='threads for the CPU arm (0 = min(16, usable cores): the tiny-op-bound reference path gets '
                         'SLOWER beyond that: 128 threads measured 40x slower than 8 on the GPU box)')
    return ap.parse_args()


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.samples, self.stop_flag, self.index = [], False, index
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.FIELDS,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(',')]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            time.sleep(0.2)

    def start(self):
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        self.thread.join(timeout=3)
        sm = sorted(int(float(s[0])) for s in self.samples if s[0].replace('.', '').isdigit())
        reasons = set()
        for s in self.samples:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), s[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        mx = [int(float(s[1])) for s in self.samples if s[1].replace('.', '').isdigit()]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(self.samples)}


# --------------------------------------------------------------------------------------------- CPU reference arm
def usable_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_threads(args):
    return args.cpu_threads if args.cpu_threads > 0 else min(16, usable_cores())


def cpu_reference_rate(batch, unroll, repeats=2, threads=None):
    """The reference's CPU PyTorch path, as restated by the oracle (kind 'port'): rl_learner_forward + loss +
    backward on a bounded sample.  Returns (frames/s, seconds per step, cores)."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import alphastar_ref as O
    from distar_b200.params import init_state_dict
    from distar_b200.synth import synth_rl_batch, tree_clone
    cores = threads or min(16, usable_cores())
    torch.set_num_threads(cores)
    sd = init_state_dict(seed=0)
    P = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in sd.items()}
    data = synth_rl_batch(batch, unroll, seed=0)

    def one():
        for p in P.values():
            p.grad = None
        info = O.rl_loss(O.rl_learner_forward(P, **tree_clone(data)))
        info['total_loss'].backward()
        return float(info['total_loss'])
    one()
    best = float('inf')
    for _ in range(repeats):
        t = time.time()
        one()
        best = min(best, time.time() - t)
    return batch * unroll / best, best, cores


def run_reference(args, rank):
    if rank != 0:
        return
    times = []
    cores = cpu_threads(args)
    # each "step" is one bounded sample (cpu_batch x cpu_unroll frames)
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import alphastar_ref as O
    from distar_b200.params import init_state_dict
    from distar_b200.synth import synth_rl_batch, tree_clone
    torch.set_num_threads(cores)
    sd = init_state_dict(seed=0)
    P = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in sd.items()}
    data = synth_rl_batch(args.cpu_batch, args.cpu_unroll, seed=0)

    def one():
        for p in P.values():
            p.grad = None
        info = O.rl_loss(O.rl_learner_forward(P, **tree_clone(data)))
        info['total_loss'].backward()
    for _ in range(max(args.warmup, 1)):
        one()
    for _ in range(args.steps):
        t = time.time()
        one()
        times.append(time.time() - t)
    ms = 1e3 * sum(times) / len(times)
    frames = args.cpu_batch * args.cpu_unroll
    value = frames / (ms / 1e3)
    sample = 'oracle port of the reference CPU path: rl_learner_forward+loss+backward on B=%d x T=%d frames/step' % (
        args.cpu_batch, args.cpu_unroll)
    print(json.dumps({
        'impl': 'reference', 'metric': 'learner samples/sec (unroll=32, 512 ent, 128^2 spatial)', 'value': value,
        'unit': 'samples/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'RL learner step, 512 entities, 128x128 spatial; CPU arm times a bounded sample',
                   'batch_per_step': args.cpu_batch, 'unroll': args.cpu_unroll},
        'cpu_baseline': {'value': value, 'unit': 'samples/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}))


# --------------------------------------------------------------------------------------------- GPU arm
def tree_bytes(tree):
    from distar_b200.synth import tree_map
    n = [0]
    tree_map(lambda t: n.__setitem__(0, n[0] + t.numel() * t.element_size()) or t, tree)
    return n[0]


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the two kernels BASELINE.json names, from the committed
# `ncu --set full --clock-control none` captures of exactly these launches (profiles/r01_summary.md)
NCU_SOURCE = 'profiles/r01_ncu_raw_{scatter,gemm}_final.csv'
NCU_DRAM_BYTES = {'scatter_connection': 70.4e6 + 2155e6, 'entity_mlp_gemm_terms3': 139.5e6 + 493.7e6}


def kernel_rooflines(dev, peaks):
    """Stand-alone CUDA-event timings of the two kernels BASELINE.json names, at the bench shapes."""
    from distar_b200 import ops
    out = {}
    # scatter_connection: N rows of the bench batch; output 2 MiB/obs >> L2, so no flush needed
    N, E = 1056, 512
    proj = torch.randn(N, E, 32, device=dev)
    ex = torch.randint(0, 128, (N, E), device=dev, dtype=torch.uint8)
    ey = torch.randint(0, 128, (N, E), device=dev, dtype=torch.uint8)
    en = torch.full((N,), E, device=dev, dtype=torch.int64)
    for _ in range(3):
        ops.scatter_connection(proj, ex, ey, en, 128, 128)
    s, e = torch.cuda.Event(True), torch.cuda.Event(True)
    reps = 10
    torch.cuda.synchronize()
    s.record()
    for _ in range(reps):
        ops.scatter_connection(proj, ex, ey, en, 128, 128)
    e.record()
    torch.cuda.synchronize()
    dt = s.elapsed_time(e) / reps / 1e3
    bytes_per_obs = 32 * 128 * 128 * 4 + E * 32 * 4 + E * 2
    ach = N * bytes_per_obs / dt / 1e9
    out['scatter_connection'] = {'bound': 'hbm', 'achieved': ach, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                                 'frac': ach / peaks['hbm_gbs'], 'traffic': NCU_DRAM_BYTES['scatter_connection'],
                                 'traffic_source': NCU_SOURCE, 'algorithmic_bytes': N * bytes_per_obs,
                                 'us_per_launch': dt * 1e6,
                                 'shape': 'N=%d obs, 512 entities, 32ch, 128x128' % N,
                                 'peak_source': peaks['source']}
    # entity-transformer MLP GEMM: [M,256] x [1024,256]^T with 3-term split, at the shape the learner step launches
    # (M = encoder_chunk 264 obs x 512 tokens) and at 4x that.  The launches are replayed from a CUDA graph so the
    # (Python + tensor-map encode) host time of a launch cannot pad the device-side duration.
    from distar_b200 import lib as _lib
    K, Nn = 256, 1024
    for M, tag in ((264 * 512, ''), (1024 * 512, '_M524288')):
        a = torch.randn(M, K, device=dev)
        w = torch.randn(Nn, K, device=dev) / 16
        b = torch.randn(Nn, device=dev)
        a_hi, a_lo = ops.split_bf16(a)
        w_hi, w_lo = ops.split_bf16(w)
        c = torch.empty(M, Nn, device=dev)
        for terms, bn, mc in ((3, 0, 1), (3, 128, 1), (1, 0, 1), (3, 256, 4), (3, 128, 4), (1, 256, 4)):
            def run():
                _lib.gemm_ex(a_hi=a_hi, a_lo=a_lo, b_hi=w_hi, b_lo=w_lo, bias=b, alpha=1.0, relu=1, terms=terms, c=c, m=M,
                             n=Nn, k=K, batch=1, inner=1, splits=1, bn=bn, mc=mc)
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for _ in range(reps):
                    run()
            graph.replay()
            torch.cuda.synchronize()
            s.record()
            graph.replay()
            e.record()
            torch.cuda.synchronize()
            dt = s.elapsed_time(e) / reps / 1e3
            flops = 2.0 * M * K * Nn * terms
            ach = flops / dt / 1e12
            key = 'entity_mlp_gemm_terms%d' % terms + ('_bn%d' % bn if bn else '') + ('_pair' if mc == 4 else '') + tag
            out[key] = {
                'bound': 'tensor', 'achieved': ach, 'peak': peaks['bf16_tflops'], 'unit': 'TFLOP/s',
                'frac': ach / peaks['bf16_tflops'], 'traffic': NCU_DRAM_BYTES.get(key), 'us_per_launch': dt * 1e6,
                'shape': 'M=%d K=%d N=%d, %d bf16 MMA terms (tensor-core flops counted), tile %sx%s' % (
                    M, K, Nn, terms, '256(CTA pair)' if mc == 4 else '128', bn if bn else 'auto(256)'),
                'peak_source': peaks['source']}
        del a, a_hi, a_lo, c
    return out


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': d['hbm_gbs'], 'bf16_tflops': d['bf16_tflops'], 'source': 'measured (MEASURED_PEAKS.json)'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'source': 'fallback (B200_PROFILING.md)'}


def run_b200(args, rank, world, local_rank):
    import torch.distributed as dist
    from distar_b200 import lib
    from distar_b200.learner import RLLearner
    from distar_b200.model import Model
    from distar_b200.synth import synth_rl_batch, tree_map
    lib.load()
    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl b200 needs a CUDA device: the product path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl')
    B, T = args.batch, args.unroll
    cfg = {'model': {'spatial_x': 128, 'spatial_y': 128, 'enable_baselines': ['winloss']}}
    model = Model(cfg, use_value_network=True, seed=0, gemm_terms=args.terms, encoder_chunk=args.encoder_chunk,
                  checkpoint_encoder=not args.no_checkpoint, keep_chunks=args.keep_chunks).cuda()
    learner = RLLearner(model, 'MP0', None, lr=1e-5, max_norm=1.0)
    host = synth_rl_batch(B, T, seed=1000 * rank)
    host = tree_map(lambda t: t.pin_memory(), host)
    h2d = tree_bytes(host)
    resident = tree_map(lambda t: t.to(dev, non_blocking=True), host)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps):
        barrier()
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record()
        for _ in range(steps):
            step_fn()
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / steps

    def step_resident():
        # the loss zeroes value[-1] in place of the model OUTPUT only; inputs are never modified
        learner._train(resident)

    last_loss = [None]
    last_info = [None]
    copy_stream = torch.cuda.Stream(device=dev)
    # two persistent device staging batches (double buffering): allocating 1.5 GB of fresh device tensors per step on a side
    # stream made the caching allocator fall back to cudaMalloc / cudaFree under a 150 GiB working set (+46 ms per step)
    stage = [tree_map(lambda t: torch.empty(t.shape, dtype=t.dtype, device=dev), host) for _ in range(2)]
    ready = [None, None]      # copy-stream event: staging batch i holds the next batch
    consumed = [None, None]   # main-stream event: the step that read staging batch i is completely queued behind it
    counter = [0]
    copy_times = []

    def _copy_tree(dst, src):
        if isinstance(src, dict):
            for k in src:
                _copy_tree(dst[k], src[k])
        elif isinstance(src, (list, tuple)):
            for d, s_ in zip(dst, src):
                _copy_tree(d, s_)
        elif torch.is_tensor(src):
            dst.copy_(src, non_blocking=True)

    def prefetch(i):
        """host -> device copy of the NEXT step's batch from pinned memory on a side stream (what the reference's
        `use_async_cuda` dataloader does, rl_dataloader.py:113-127); every step pays for exactly one such copy."""
        with torch.cuda.stream(copy_stream):
            if consumed[i] is not None:
                copy_stream.wait_event(consumed[i])        # the step that used this staging batch must be done with it
            t0 = torch.cuda.Event(enable_timing=True) if os.environ.get('DSB_E2E_TIMECOPY') == '1' else None
            if t0 is not None:
                t0.record(copy_stream)
            _copy_tree(stage[i], host)
            ready[i] = torch.cuda.Event(enable_timing=t0 is not None)
            ready[i].record(copy_stream)
            if t0 is not None:
                copy_times.append((t0, ready[i]))

    def step_e2e():
        if os.environ.get('DSB_E2E_NOCOPY') == '1':        # DEV ONLY: isolates the cost of the per-step result read
            info = learner._train(resident)
            last_loss[0] = info['_total_loss_value']
            last_info[0] = info
            return
        i = counter[0] & 1
        counter[0] += 1
        if ready[i] is None:
            prefetch(i)
        torch.cuda.current_stream().wait_event(ready[i])
        ready[i] = None
        prefetch(1 - i)                                     # next step's batch; overlaps with this step's compute
        info = learner._train(stage[i])
        consumed[i] = torch.cuda.Event()
        consumed[i].record(torch.cuda.current_stream())
        # device -> host read of the step result: the loss and the ~45 logged scalars arrive in ONE asynchronous copy queued
        # right after the loss (rl_loss.LazyScalars); reading them waits for forward + loss only, so the host queues the next
        # step while this step's backward is still running
        last_loss[0] = info['_total_loss_value']
        last_info[0] = info

    if os.environ.get('DSB_ANOMALY') == '1':
        torch.autograd.set_detect_anomaly(True)
    for _ in range(args.warmup):
        step_resident()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    l0 = lib.launch_count()
    ms = timed(step_resident, args.steps)
    launches = lib.launch_count() - l0
    clk = clocks.stop() if rank == 0 else None
    e2e = None
    if not args.no_e2e:
        step_e2e()
        ms_e2e = timed(step_e2e, args.steps)
        if copy_times:
            torch.cuda.synchronize()
            print('H2D batch copy durations (ms):', ['%.1f' % a.elapsed_time(b) for a, b in copy_times], file=sys.stderr)
        e2e = {'value': world * B * T / (ms_e2e / 1e3), 'unit': 'samples/s', 'h2d_bytes_per_step': h2d,
               'd2h_bytes_per_step': 4 * len(last_info[0]) if last_info[0] is not None else 0, 'ms_per_step': ms_e2e}
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # secondary figure (BASELINE configs[1]): actor-side inference, batch 32, forward + sampling, per GPU
    from distar_b200.synth import synth_obs
    obs = tree_map(lambda t: t.to(dev), synth_obs(32, seed=7))
    with torch.no_grad():
        for _ in range(2):
            model.compute_logp_action(**obs)
        torch.cuda.synchronize()
        i0, i1 = torch.cuda.Event(True), torch.cuda.Event(True)
        i0.record()
        for _ in range(5):
            model.compute_logp_action(**obs)
        i1.record()
        torch.cuda.synchronize()
    infer_ms = i0.elapsed_time(i1) / 5
    peaks = load_peaks()
    roofs = kernel_rooflines(dev, peaks)
    line = {
        'metric': 'learner samples/sec (unroll=32, 512 ent, 128^2 spatial)', 'value': world * B * T / (ms / 1e3),
        'unit': 'samples/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32 (bf16x3 split products on tcgen05, fp32 accumulate)' if args.terms == 3 else 'bf16',
        'data': 'synthetic',
        'config': {'workload': 'RL learner step (rl_learner_forward + V-trace/UPGO/TD/entropy/KL loss + backward + '
                               'clip + Adam), BASELINE configs[3] per rank',
                   'batch_per_gpu': B, 'unroll': T, 'entities': 512, 'spatial': '128x128', 'global_batch': world * B,
                   'parallelism': 'dp%d' % world, 'encoder_chunk': args.encoder_chunk,
                   'entity_chunks_recomputed_in_backward': max(0, -(-(T + 1) * B // args.encoder_chunk) - args.keep_chunks),
                   'l2': 'inputs and activations (GBs per step) far exceed the 126 MB L2; no flush needed'},
        'clocks': clk, 'e2e': e2e, 'gpu_launches': int(launches),
        'peak_hbm_gib': round(torch.cuda.max_memory_allocated() / 2 ** 30, 1),
        'inference': {'workload': 'compute_logp_action, batch 32 (forward + sampling), 1 GPU', 'ms_per_call': infer_ms,
                      'obs_per_s': 32 / (infer_ms / 1e3)},
        'roofline': roofs['entity_mlp_gemm_terms3'] if args.terms == 3 else roofs['entity_mlp_gemm_terms1'],
        'rooflines': roofs,
    }
    if not args.no_cpu_baseline:
        v, sec, cores = cpu_reference_rate(args.cpu_batch, args.cpu_unroll, threads=cpu_threads(args))
        line['cpu_baseline'] = {'value': v, 'unit': 'samples/s', 'cores': cores, 'kind': 'port',
                                'sample': 'oracle port, rl_learner_forward+loss+backward, B=%d x T=%d frames, best of 2 '
                                          '(%.1f s/step)' % (args.cpu_batch, args.cpu_unroll, sec)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
