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
    ap.add_argument('--workload', default='rl', choices=['rl', 'sl', 'league'],
                    help='rl = BASELINE configs[3] (the metric; default).  sl = configs[2]: SLLearner step, batch 64 x unroll 32. '
                         'league = configs[4]: main / main-exploiter / league-exploiter players training side by side on '
                         'disjoint GPU groups (4/2/2 of 8), one NCCL communicator per player')
    ap.add_argument('--batch', type=int, default=0, help='trajectories per rank (BASELINE: 128 for rl / league, 64 for sl)')
    ap.add_argument('--unroll', type=int, default=32, help='unroll length (BASELINE: 32)')
    ap.add_argument('--encoder-chunk', type=int, default=264)
    ap.add_argument('--terms', type=int, default=3, help='tensor-core products per GEMM: 3 = fp32-class (parity), 1 = bf16')
    ap.add_argument('--no-checkpoint', action='store_true', help='keep encoder activations instead of recomputing them')
    ap.add_argument('--keep-chunks', type=int, default=16,
                    help='number of encoder chunks whose entity-transformer activations are kept (not recomputed)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--value-feature', action='store_true',
                    help='rl / league: learner.use_value_feature True (the reference self-play default): ValueEncoder in front of the '
                         'baselines, value_feature in the batch.  Off by default: BASELINE configs do not include it')
    ap.add_argument('--e2e-format', default='compact', choices=['compact', 'padded'],
                    help='what crosses PCIe every step of the e2e measurement: compact = un-padded trajectories expanded on the '
                         'GPU (distar_b200.batch, the product path); padded = the reference collate layout (1.47 GB / step)')
    ap.add_argument('--cpu-batch', type=int, default=4)
    ap.add_argument('--cpu-unroll', type=int, default=8)
    ap.add_argument('--cpu-all-cores', action='store_true',
                    help='also time the CPU arm on ALL host cores (BASELINE.md section 4 planned that; it is ~40x SLOWER than 16 '
                         'threads on the GPU box, about a minute per pass, so it is opt-in; the recorded figure is in profiles/)')
    ap.add_argument('--cpu-threads', type=int, default=0,
                    help='threads for the CPU arm (0 = min(16, usable cores): the tiny-op-bound reference path gets '
                         'SLOWER beyond that: 128 threads measured 40x slower than 8 on the GPU box)')
    args = ap.parse_args()
    if args.batch <= 0:
        args.batch = 64 if args.workload == 'sl' else 128
    return args


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


def _oracle_step(batch, unroll, sl=False):
    """One learner step of the reference's CPU PyTorch path as restated by the oracle (kind 'port'): forward + loss +
    backward on `batch` x `unroll` frames.  The ONLY place outside tests/ and smoke() where bench.py executes oracle/ code:
    as the thing being timed for the CPU baseline, never on the product path."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import alphastar_ref as O
    from distar_b200.params import init_state_dict
    from distar_b200.synth import synth_rl_batch, synth_sl_batch, tree_clone
    sd = init_state_dict(seed=0)
    P = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in sd.items()}
    if sl:
        data = synth_sl_batch(batch, unroll, seed=0)
        hidden = [(torch.zeros(batch, 384), torch.zeros(batch, 384)) for _ in range(3)]
        fwd = {k: data[k] for k in ('spatial_info', 'entity_info', 'scalar_info', 'entity_num', 'selected_units_num',
                                    'traj_lens', 'action_info')}

        def one():
            for p in P.values():
                p.grad = None
            logits, _, _ = O.sl_train(P, **tree_clone(fwd), hidden_state=tree_clone(hidden))
            info = O.sl_loss(logits, data['action_info'], data['action_mask'], data['selected_units_num'])
            info['total_loss'].backward()
            return float(info['total_loss'])
    else:
        data = synth_rl_batch(batch, unroll, seed=0)

        def one():
            for p in P.values():
                p.grad = None
            info = O.rl_loss(O.rl_learner_forward(P, **tree_clone(data)))
            info['total_loss'].backward()
            return float(info['total_loss'])
    return one


def cpu_reference_rate(batch, unroll, repeats=2, threads=None, sl=False, warm=True):
    """Returns (frames/s, seconds per step, cores) of the oracle port on a bounded sample."""
    cores = threads or min(16, usable_cores())
    torch.set_num_threads(cores)
    one = _oracle_step(batch, unroll, sl)
    if warm:
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
    torch.set_num_threads(cores)
    sl = args.workload == 'sl'
    one = _oracle_step(args.cpu_batch, args.cpu_unroll, sl)     # each "step" is one bounded sample (cpu_batch x cpu_unroll frames)
    for _ in range(max(args.warmup, 1)):
        one()
    for _ in range(args.steps):
        t = time.time()
        one()
        times.append(time.time() - t)
    ms = 1e3 * sum(times) / len(times)
    frames = args.cpu_batch * args.cpu_unroll
    value = frames / (ms / 1e3)
    sample = 'oracle port of the reference CPU path: %s on B=%d x T=%d frames/step' % (
        'sl_train+loss+backward' if sl else 'rl_learner_forward+loss+backward', args.cpu_batch, args.cpu_unroll)
    print(json.dumps({
        'impl': 'reference', 'metric': 'learner samples/sec (unroll=32, 512 ent, 128^2 spatial)', 'value': value,
        'unit': 'samples/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': '%s learner step, 512 entities, 128x128 spatial; CPU arm times a bounded sample' % (
            'SL' if sl else 'RL'), 'batch_per_step': args.cpu_batch, 'unroll': args.cpu_unroll},
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
NCU_SOURCE = 'profiles/r02_ncu_raw_{scatter,gemm}.csv (the GEMM capture is the single-CTA variant <256,1> of this launch: same operands and output bytes as the CTA-pair variant timed here)'
NCU_DRAM_BYTES = {'scatter_connection': 70.4e6 + 2155.0e6, 'entity_mlp_gemm_terms3': 139.6e6 + 493.7e6}


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
        for terms, bn, mc in ((3, 0, 0), (3, 256, 1), (3, 128, 1), (1, 0, 1), (1, 256, 4)):
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
            # `achieved` / `frac` count the ALGORITHMIC flop of the launch (2*M*K*N); the bf16 split forms every
            # product `terms` times on the tensor cores: that tensor work is reported beside it
            flops = 2.0 * M * K * Nn
            ach = flops / dt / 1e12
            key = 'entity_mlp_gemm_terms%d' % terms + ('_bn%d' % bn if bn else '') + ('_pair' if mc == 4 else '') + \
                ('_single' if (bn, mc) == (256, 1) else '') + tag
            out[key] = {
                'bound': 'tensor', 'achieved': ach, 'peak': peaks['bf16_tflops'], 'unit': 'TFLOP/s',
                'frac': ach / peaks['bf16_tflops'], 'achieved_tensor_work': ach * terms,
                'frac_tensor_work': ach * terms / peaks['bf16_tflops'], 'traffic': NCU_DRAM_BYTES.get(key),
                'traffic_source': NCU_SOURCE if NCU_DRAM_BYTES.get(key) else None, 'us_per_launch': dt * 1e6,
                'algorithmic_flop': flops,
                'shape': 'M=%d K=%d N=%d, %d bf16 MMA products per element, tile %sx%s' % (
                    M, K, Nn, terms, '256(CTA pair, cta_group::2)' if mc == 4 or (mc == 0 and terms == 3) else '128',
                    bn if bn else 'auto(256)'),
                'peak_source': peaks['source']}
        del a, a_hi, a_lo, c
    # the down-projection of the same MLP: [M,1024] x [256,1024]^T (16 k blocks per tile: the mainloop-bound shape)
    M, K, Nn = 264 * 512, 1024, 256
    a = torch.randn(M, K, device=dev)
    w = torch.randn(Nn, K, device=dev) / 32
    b = torch.randn(Nn, device=dev)
    a_hi, a_lo = ops.split_bf16(a)
    w_hi, w_lo = ops.split_bf16(w)
    c = torch.empty(M, Nn, device=dev)

    def run2():
        _lib.gemm_ex(a_hi=a_hi, a_lo=a_lo, b_hi=w_hi, b_lo=w_lo, bias=b, alpha=1.0, relu=1, terms=3, c=c, m=M, n=Nn, k=K,
                     batch=1, inner=1, splits=1)
    for _ in range(3):
        run2()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(reps):
            run2()
    graph.replay()
    torch.cuda.synchronize()
    s.record()
    graph.replay()
    e.record()
    torch.cuda.synchronize()
    dt = s.elapsed_time(e) / reps / 1e3
    flops = 2.0 * M * K * Nn
    ach = flops / dt / 1e12
    out['entity_ffn2_gemm_terms3'] = {
        'bound': 'tensor', 'achieved': ach, 'peak': peaks['bf16_tflops'], 'unit': 'TFLOP/s', 'frac': ach / peaks['bf16_tflops'],
        'achieved_tensor_work': 3 * ach, 'frac_tensor_work': 3 * ach / peaks['bf16_tflops'], 'traffic': None,
        'us_per_launch': dt * 1e6, 'algorithmic_flop': flops,
        'shape': 'M=%d K=%d N=%d, 3 bf16 MMA products per element' % (M, K, Nn), 'peak_source': peaks['source']}
    return out


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': d['hbm_gbs'], 'bf16_tflops': d['bf16_tflops'],
                'bf16_tflops_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']), 'source': 'measured (MEASURED_PEAKS.json)'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1590.0, 'source': 'fallback (B200_PROFILING.md)'}


def gemm_family_in_step(step_fn):
    """One extra learner step with CUDA events around EVERY tcgen05 GEMM launch (distar_b200.lib.GEMM_TRACE): the time the
    dominant kernel family takes inside the step, its algorithmic flop (2*m*n*k per launch) and the tensor work
    actually issued (x3 / x2 products of the bf16 split).  The events perturb the step a little, so this step is not part of
    the timed region; the share it reports is checked against the ncu launch list under profiles/."""
    from distar_b200 import lib as _lib
    torch.cuda.synchronize()
    _lib.GEMM_TRACE = []
    try:
        step_fn()
        torch.cuda.synchronize()
        trace = _lib.GEMM_TRACE
    finally:
        _lib.GEMM_TRACE = None
    ms = sum(a.elapsed_time(b) for a, b, _, _ in trace)
    flops = sum(f for _, _, f, _ in trace)
    work = sum(f * p for _, _, f, p in trace)
    return {'bound': 'tensor', 'launches': len(trace), 'ms_per_step': ms, 'algorithmic_tflop': flops / 1e12,
            'achieved': flops / ms / 1e9, 'achieved_tensor_work': work / ms / 1e9, 'unit': 'TFLOP/s',
            'note': 'CUDA events around each launch; includes the launch gaps the events themselves introduce'}


LEAGUE_PLAYERS = ('MP0', 'ME0', 'EP0')        # main player, main exploiter, league exploiter (league/player.py)


def league_layout(world):
    """BASELINE configs[4]: the three player kinds train side by side, each on its own GPUs with its own gradient exchange
    (rl_train.py:26-51 starts one learner job per player id).  8 GPUs -> MP0 on 4, ME0 on 2, EP0 on 2."""
    if world >= 8:
        sizes = [world // 2, world // 4, world - world // 2 - world // 4]
    elif world >= 4:
        sizes = [world - 2, 1, 1]
    elif world == 3:
        sizes = [1, 1, 1]
    elif world == 2:
        sizes = [1, 1, 0]
    else:
        sizes = [1, 0, 0]
    groups, r = [], 0
    for pid, n in zip(LEAGUE_PLAYERS, sizes):
        if n:
            groups.append((pid, list(range(r, r + n))))
            r += n
    return groups


def run_b200(args, rank, world, local_rank):
    import torch.distributed as dist
    from distar_b200 import lib
    from distar_b200.learner import RLLearner, SLLearner
    from distar_b200.model import Model
    from distar_b200.rl_loss import USER_LEARNER_CFG
    from distar_b200.synth import synth_rl_batch, synth_sl_batch, tree_map
    lib.load()
    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl b200 needs a CUDA device: the product path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl')
    B, T = args.batch, args.unroll
    rl = args.workload != 'sl'
    cfg = {'model': {'spatial_x': 128, 'spatial_y': 128, 'enable_baselines': ['winloss']},
           'learner': {'use_value_feature': bool(args.value_feature and rl)}}
    model = Model(cfg, use_value_network=rl, seed=0, gemm_terms=args.terms, encoder_chunk=args.encoder_chunk,
                  checkpoint_encoder=not args.no_checkpoint, keep_chunks=args.keep_chunks).cuda()
    player, group, group_ranks = 'MP0', None, list(range(world))
    layout = None
    if args.workload == 'league':
        layout = league_layout(world)
        for pid, ranks in layout:                      # every rank creates every group (torch.distributed contract)
            g = dist.new_group(ranks) if world > 1 else None
            if rank in ranks:
                player, group, group_ranks = pid, g, ranks
        # the exploiters of a league run with DAPO off (rl_loss.py:22-24); MP keeps the user config (use_dapo False by default)
        learner = RLLearner(model, player, dict(USER_LEARNER_CFG), lr=1e-5, max_norm=1.0, group=group)
        host = synth_rl_batch(B, T, seed=1000 * rank + 17 * LEAGUE_PLAYERS.index(player), value_feature=args.value_feature)
    elif rl:
        learner = RLLearner(model, 'MP0', None, lr=1e-5, max_norm=1.0)
        host = synth_rl_batch(B, T, seed=1000 * rank, value_feature=args.value_feature)
    else:
        # bin/sl_user_config.yaml: Adam(lr 1e-3, weight_decay 1e-5), clip 'momentum_norm' 1.4, su_mask off, warm-up 20000
        sl_cfg = {'learner': {'su_mask': False, 'learning_rate': 1e-3, 'weight_decay': 1e-5, 'use_warmup': True,
                              'warm_up_steps': 20000, 'grad_clip': {'type': 'momentum_norm', 'threshold': 1.4},
                              'data': {'batch_size': B}}}
        learner = SLLearner(model, sl_cfg, ignore_steps=-1)       # the reference's 6 no-update iterations are start-up only
        host = synth_sl_batch(B, T, seed=1000 * rank)
    compact = rl and args.e2e_format == 'compact'
    padded_bytes = tree_bytes(host)
    if compact:
        from distar_b200.batch import compact_rl_batch, expand_rl_batch
        wire = tree_map(lambda t: t.pin_memory(), compact_rl_batch(host))      # what an actor-side sender would ship
    host = tree_map(lambda t: t.pin_memory(), host)
    if not compact:
        wire = host
    h2d = tree_bytes(wire)
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
        mine = s.elapsed_time(e) / steps
        ms = torch.tensor([mine], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), mine

    def loss_value(info):
        return info['_total_loss_value'] if rl else info['total_loss']

    def step_resident():
        # the loss zeroes value[-1] in place of the model OUTPUT only; inputs are never modified
        learner._train(resident)

    last_loss = [None]
    last_info = [None]
    copy_stream = torch.cuda.Stream(device=dev)
    # two persistent device staging batches (double buffering): allocating 1.5 GB of fresh device tensors per step on a side
    # stream made the caching allocator fall back to cudaMalloc / cudaFree under a 150 GiB working set (+46 ms per step)
    stage = [tree_map(lambda t: torch.empty(t.shape, dtype=t.dtype, device=dev), wire) for _ in range(2)]
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
            _copy_tree(stage[i], wire)
            ready[i] = torch.cuda.Event(enable_timing=t0 is not None)
            ready[i].record(copy_stream)
            if t0 is not None:
                copy_times.append((t0, ready[i]))

    def step_e2e():
        if os.environ.get('DSB_E2E_NOCOPY') == '1':        # DEV ONLY: isolates the cost of the per-step result read
            info = learner._train(resident)
            last_loss[0] = loss_value(info)
            last_info[0] = info
            return
        i = counter[0] & 1
        counter[0] += 1
        if ready[i] is None:
            prefetch(i)
        torch.cuda.current_stream().wait_event(ready[i])
        ready[i] = None
        prefetch(1 - i)                                     # next step's batch; overlaps with this step's compute
        # compact wire format: the padded reference layout is assembled in HBM by three kernels (inside the timed region)
        info = learner._train(expand_rl_batch(wire, staged=stage[i]) if compact else stage[i])
        consumed[i] = torch.cuda.Event()
        consumed[i].record(torch.cuda.current_stream())
        # device -> host read of the step result: the loss and the ~45 logged scalars arrive in ONE asynchronous copy queued
        # right after the loss (rl_loss.LazyScalars); reading them waits for forward + loss only, so the host queues the next
        # step while this step's backward is still running
        last_loss[0] = loss_value(info)
        last_info[0] = info

    if os.environ.get('DSB_ANOMALY') == '1':
        torch.autograd.set_detect_anomaly(True)
    for _ in range(args.warmup):
        step_resident()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    l0 = lib.launch_count()
    ms, ms_mine = timed(step_resident, args.steps)
    launches = lib.launch_count() - l0
    # the dominant kernel family, timed live inside the step: CUDA events around every tcgen05 GEMM launch of ONE extra step.
    # EVERY rank runs that step (it contains the gradient all-reduce of the rank's group: a step on rank 0 alone would wait
    # for its peers forever); rank 0's trace is the one reported.
    gemm_in_step = gemm_family_in_step(step_resident) if os.environ.get('DSB_NO_GEMM_TIMING') != '1' else None
    clk = clocks.stop() if rank == 0 else None
    e2e = None
    ms_e2e_mine = None
    if not args.no_e2e:
        step_e2e()
        ms_e2e, ms_e2e_mine = timed(step_e2e, args.steps)
        if copy_times:
            torch.cuda.synchronize()
            print('H2D batch copy durations (ms):', ['%.1f' % a.elapsed_time(b) for a, b in copy_times], file=sys.stderr)
        n_scalars = len(last_info[0]) if last_info[0] is not None else 0
        e2e = {'value': world * B * T / (ms_e2e / 1e3), 'unit': 'samples/s', 'h2d_bytes_per_step': h2d,
               'd2h_bytes_per_step': 4 * n_scalars, 'ms_per_step': ms_e2e,
               'wire_format': 'compact trajectories, expanded on the GPU (distar_b200.batch)' if compact else
                              'padded reference collate layout', 'padded_layout_bytes': padded_bytes}
    per_player = None
    if args.workload == 'league':
        # every player's learners report their own step time (max within the group); a player's rate = its GPUs * B * T / that
        mine = torch.tensor([ms_mine, ms_e2e_mine or 0.0], device=dev)
        allv = [torch.zeros_like(mine) for _ in range(world)] if world > 1 else [mine]
        if world > 1:
            dist.all_gather(allv, mine)
        per_player = {}
        for pid, ranks in layout:
            t_res = max(float(allv[r][0]) for r in ranks)
            t_e2e = max(float(allv[r][1]) for r in ranks)
            per_player[pid] = {'gpus': len(ranks), 'ms_per_step': t_res, 'samples_per_s': len(ranks) * B * T / (t_res / 1e3),
                               'e2e_samples_per_s': (len(ranks) * B * T / (t_e2e / 1e3)) if t_e2e else None}
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
    # the same call served from a captured CUDA graph (distar_b200.serving.InferenceServer): host batch in pinned memory ->
    # static device buffers -> replay (64 fixed pointer steps) -> results; timed end to end per request
    serve = None
    try:
        from distar_b200.serving import InferenceServer
        host_obs = tree_map(lambda t: t.pin_memory(), synth_obs(32, seed=7))
        server = InferenceServer(model, host_obs, su_steps=64)
        for _ in range(2):
            server.infer(host_obs)
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(10):
            server.infer(host_obs)
        torch.cuda.synchronize()
        serve_ms = (time.time() - t0) * 1e3 / 10
        serve = {'ms_per_request': serve_ms, 'obs_per_s': 32 / (serve_ms / 1e3), 'batch': 32,
                 'note': 'CUDA-graph replay incl. H2D of the observation batch and the host read of the result'}
        del server
    except Exception as exc:                                    # the learner metric must not depend on the serving extra
        serve = {'error': repr(exc)[:200]}
    peaks = load_peaks()
    roofs = kernel_rooflines(dev, peaks)
    workload = {
        'rl': 'RL learner step (rl_learner_forward + V-trace/UPGO/TD/entropy/KL loss + backward + clip + Adam), BASELINE '
              'configs[3] per rank',
        'sl': 'SL learner step (sl_train + 6 masked cross-entropies + backward + Adam(weight decay)), BASELINE configs[2]',
        'league': 'league step: %s training side by side (one RL learner step each, own NCCL communicator per player), '
                  'BASELINE configs[4]' % ' / '.join('%s x%d' % (p, len(r)) for p, r in (layout or []))}[args.workload]
    main = dict(roofs['entity_mlp_gemm_terms3'] if args.terms == 3 else roofs['entity_mlp_gemm_terms1'])
    main['kernels'] = {'scatter_connection': roofs['scatter_connection'], 'entity_ffn1_gemm': dict(main),
                       'entity_ffn2_gemm': roofs.get('entity_ffn2_gemm_terms3')}
    if gemm_in_step is not None:
        gemm_in_step['peak'] = peaks['bf16_tflops_sustained']
        gemm_in_step['frac_tensor_work'] = gemm_in_step['achieved_tensor_work'] / peaks['bf16_tflops_sustained']
        gemm_in_step['share_of_step'] = gemm_in_step['ms_per_step'] / ms
        main['kernels']['tcgen05_gemm_family_in_step'] = gemm_in_step
    line = {
        'metric': 'learner samples/sec (unroll=32, 512 ent, 128^2 spatial)', 'value': world * B * T / (ms / 1e3),
        'unit': 'samples/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32 (bf16x3 split products on tcgen05, fp32 accumulate)' if args.terms == 3 else 'bf16',
        'data': 'synthetic',
        'config': {'workload': workload,
                   'batch_per_gpu': B, 'unroll': T, 'entities': 512, 'spatial': '128x128', 'global_batch': world * B,
                   'use_value_feature': bool(args.value_feature and rl),
                   'parallelism': 'dp%d' % world if args.workload != 'league' else
                                  '+'.join('dp%d(%s)' % (len(r), p) for p, r in layout),
                   'encoder_chunk': args.encoder_chunk,
                   'entity_chunks_recomputed_in_backward': max(0, -(-(T + (1 if rl else 0)) * B // args.encoder_chunk) - args.keep_chunks),
                   'l2': 'inputs and activations (GBs per step) far exceed the 126 MB L2; no flush needed'},
        'clocks': clk, 'e2e': e2e, 'gpu_launches': int(launches),
        'peak_hbm_gib': round(torch.cuda.max_memory_allocated() / 2 ** 30, 1),
        'inference': {'workload': 'compute_logp_action, batch 32 (forward + sampling), 1 GPU', 'ms_per_call': infer_ms,
                      'obs_per_s': 32 / (infer_ms / 1e3), 'graph_server': serve},
        'roofline': main,
        'rooflines': roofs,
    }
    if per_player is not None:
        line['players'] = per_player
    if not args.no_cpu_baseline:
        v, sec, cores = cpu_reference_rate(args.cpu_batch, args.cpu_unroll, threads=cpu_threads(args), sl=not rl)
        line['cpu_baseline'] = {'value': v, 'unit': 'samples/s', 'cores': cores, 'kind': 'port',
                                'sample': 'oracle port, %s, B=%d x T=%d frames, best of 2 (%.1f s/step)' % (
                                    'rl_learner_forward+loss+backward' if rl else 'sl_train+loss+backward',
                                    args.cpu_batch, args.cpu_unroll, sec)}
        if usable_cores() > cores and args.cpu_all_cores:
            # BASELINE.md planned the CPU arm on all host cores; it is SLOWER there (tiny-op bound), so both are on record
            v2, sec2, cores2 = cpu_reference_rate(args.cpu_batch, args.cpu_unroll, repeats=1, threads=usable_cores(), sl=not rl,
                                                   warm=False)     # one pass only: it is ~40x slower than the 16-thread run
            line['cpu_baseline']['all_cores'] = {'value': v2, 'cores': cores2, 's_per_step': sec2}
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
