"""Benchmark of the B200-native hot paths of zeis/deepbelief (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[4], the configuration the CD metric is quoted on): synthetic binary RBM
V=4096 x H=8192, GLOBAL batch 16384 rows of Bernoulli(0.5) data, CD-10, Adam(1e-3). A step is one full
CD-10 iteration: positive phase, 10 Gibbs steps, dW/db/dc, (N>1: NCCL sum-all-reduce of the packed
gradient), optimizer update and split-weight refresh. The global batch is row-sharded over the ranks
(N_local = 16384 / N), i.e. strong scaling, exactly as the config names it.

Printed JSON line (rank 0): metric `RBM CD-k sample-steps/s` = N_global * k * steps / seconds, with
  value      inputs resident in HBM (a pool of device batches larger than L2), device-timed, max over ranks
  e2e        the same iterations through the layer API `RBM.train(...)` fed from pinned HOST memory, with
             the host->device copy of every batch and a device->host read of a per-step monitor inside
             the timed region
  roofline   tensor-bound: algorithmic FLOP (2k+3)*2*N*V*H per step / time of the CD chain + dW launches,
             timed live with CUDA events around that call, against MEASURED_PEAKS.json
  cpu_baseline  the numpy oracle port of the same step on the host cores, bounded sample (N=1 only)
  gplvm      second half of BASELINE.json's metric: GPLVM LML+grad evals/s at N=5000, D=784, Q=2
`--impl reference` times the CPU restatement of the reference (oracle port; TensorFlow 1.8 cannot be
installed here) on the host cores for the same metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

V, H, N_GLOBAL, K_GIBBS = 4096, 8192, 16384, 10
GP_N, GP_D, GP_Q = 5000, 784, 2
METRIC = 'RBM CD-k sample-steps/s'
UNIT = 'sample-steps/s'
CONFIG = {'workload': 'synthetic binary RBM 4096x8192, global batch 16384, CD-10, Adam(1e-3) (BASELINE configs[4])',
          'V': V, 'H': H, 'global_batch': N_GLOBAL, 'k': K_GIBBS, 'optimizer': 'adam',
          'l2_policy': 'inputs larger than L2: each device batch is %d MB (>126 MB L2), pool of 2 cycled'
                       % (N_GLOBAL * V * 4 // 2 ** 20)}


def cd_flops(n, k=K_GIBBS):
    return (2 * k + 3) * 2.0 * n * V * H


# ---------------------------------------------------------------------------------------------------------
# clocks sampling (recipe line of B200_PROFILING.md), run in a thread during the timed regions
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, str(gpu_index)

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', self.gpu, '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(',')]))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()

    def summary(self, windows):
        """median SM clock over the samples that fall inside the timed windows, and every reason seen active"""
        sm, mx, reasons, power = [], 0.0, set(), 0.0
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for t, c in self.rows:
            if len(c) < 9 or not any(a <= t <= b for a, b in windows):
                continue
            try:
                sm.append(float(c[1])); mx = max(mx, float(c[2])); power = max(power, float(c[3]))
            except ValueError:
                continue
            for name, val in zip(names, c[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx or None, 'reasons': sorted(reasons),
                'samples': len(sm), 'power_w_max': power or None}


def count_kernel_launches(fn):
    """Kernels launched on the device by fn() (one untimed call under the CUPTI-backed torch profiler), or None."""
    try:
        import torch
        from torch.profiler import profile, ProfilerActivity
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        names = [e.name for e in prof.events() if getattr(e, 'device_type', None) is not None and 'CUDA' in str(e.device_type)
                 and not e.name.lower().startswith(('memcpy', 'memset'))]
        return len(names), sorted(set(names))
    except Exception:
        return None


def ncu_dram_bytes(csv_name, kernel_substr, nth=0):
    """dram__bytes_read.sum + dram__bytes_write.sum of the nth captured launch of a kernel in a committed `ncu --page raw --csv` file"""
    import csv
    path = os.path.join(ROOT, 'profiles', csv_name)
    if not os.path.exists(path):
        return None
    mult = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    with open(path, newline='') as f:
        rows = list(csv.reader(f))
    head, units = rows[0], rows[1]
    try:
        ir, iw, ik = head.index('dram__bytes_read.sum'), head.index('dram__bytes_write.sum'), head.index('Kernel Name')
    except ValueError:
        return None
    seen = 0
    for r in rows[2:]:
        if kernel_substr in r[ik]:
            if seen == nth:
                return float(r[ir].replace(',', '')) * mult.get(units[ir], 1.0) + float(r[iw].replace(',', '')) * mult.get(units[iw], 1.0)
            seen += 1
    return None


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {'hbm_gbs': p['hbm_gbs'], 'tensor_burst': p['bf16_tflops'],
                'tensor_sustained': p.get('bf16_tflops_sustained', p['bf16_tflops']), 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'tensor_burst': 1590.0, 'tensor_sustained': 1400.0, 'source': 'fallback'}


# ---------------------------------------------------------------------------------------------------------
# CPU side (SURVEY.md 8d "CPU side-by-side"): the reference's step restated with torch-CPU ops on all host threads
# (oracle/torch_port.py, pinned to the numpy oracle by tests/test_oracle_torch_port.py); the numpy oracle stays as a
# second, slower point of comparison. TensorFlow 1.8 is not installable here (DESIGN.md "Oracle").
# ---------------------------------------------------------------------------------------------------------
CPU_ROWS = 512          # rows of the 16384-row batch one CPU step processes (bounded sample; stated in every CPU line)


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must use every core it can (BLAS pools read the
    variables at import, torch takes the explicit call)."""
    n = host_threads()
    for var in ('OMP_NUM_THREADS', 'MKL_NUM_THREADS', 'OPENBLAS_NUM_THREADS'):
        os.environ[var] = str(n)
    import torch
    torch.set_num_threads(n)
    return n


def cpu_cd_steps_torch(n_rows, steps, warmup):
    """CD-10 + Adam on `n_rows` rows of the workload: torch.matmul (MKL sgemm), sigmoid, torch.rand draws inside the step
    (the counterpart of tf.random_uniform in the graph, rbm.py:202,230), closed-form gradient, Adam. -> (s/step, threads)"""
    import torch
    from oracle import torch_port
    threads = use_all_host_threads()
    g = torch.Generator().manual_seed(0)
    port = torch_port.CdStepTorch(torch.randn(V, H, generator=g) * 0.01, torch.zeros(V), torch.zeros(H), lr=1e-3)
    v0 = (torch.rand(n_rows, V, generator=g) < 0.5).float()
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        port.step(v0, K_GIBBS, generator=g)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times), threads


def cpu_cd_steps_numpy(n_rows, steps, warmup):
    """The same step with the numpy oracle (oracle/rbm.py, fp32, BLAS threads). -> (s/step, threads)"""
    threads = use_all_host_threads()
    import numpy as np
    from oracle import rbm as orbm
    rs = np.random.RandomState(0)
    W = (rs.randn(V, H) * 0.01).astype(np.float32)
    model = orbm.RBMOracle(W, np.zeros(V, np.float32), np.zeros(H, np.float32), kind='rbm', dtype=np.float32)
    opt = orbm.Adam(1e-3)
    v0 = (rs.rand(n_rows, V) < 0.5).astype(np.float32)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        draw_h0 = rs.random_sample((n_rows, H)).astype(np.float32)
        chain = [(rs.random_sample((n_rows, V)).astype(np.float32), rs.random_sample((n_rows, H)).astype(np.float32))
                 for _ in range(K_GIBBS)]
        out = orbm.cd_k(model, v0, K_GIBBS, draw_h0, chain)
        theta = opt.step(model.pack(), out['grad'].astype(np.float32))
        model.W = theta[:V * H].reshape(V, H)
        model.b = theta[V * H:V * H + V].reshape(V, 1)
        model.c = theta[V * H + V:].reshape(1, H)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times), threads


def cpu_gplvm_evals(evals, warmup, n=GP_N):
    """GPLVM loss + gradient at N = n, D = 784, Q = 2 in fp32 with torch-CPU ops: expanded-distance Gram, linalg.cholesky,
    solve_triangular, autograd (what TF differentiated, gplvm.py:210-244, 425-428). -> (s/eval, threads)"""
    import numpy as np
    from oracle import torch_port
    threads = use_all_host_threads()
    rs = np.random.RandomState(0)
    X = rs.normal(size=(n, GP_Q)).astype(np.float32)
    Y = (rs.uniform(size=(n, GP_D)) < 0.13).astype(np.float32)
    Y -= Y.mean(axis=0, keepdims=True)
    isp = float(np.log(np.expm1(1.0)))
    times = []
    for it in range(warmup + evals):
        t0 = time.perf_counter()
        torch_port.gplvm_loss_and_grad(X, Y, isp, isp, isp, jitter=1e-6)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times), threads


def cd_cpu_sample_text(n_rows):
    return ('CD-10 + Adam(1e-3) on %d of the %d rows of the 4096x8192 workload per step (the 33.6 M-parameter Adam update is '
            'paid in full every step)' % (n_rows, N_GLOBAL))


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    n_rows = CPU_ROWS
    sec, cores = cpu_cd_steps_torch(n_rows, args.steps, args.warmup)
    value = n_rows * K_GIBBS / sec
    sample = cd_cpu_sample_text(n_rows) + '; torch-CPU fp32 restatement of the reference graph (oracle/torch_port.py), all host threads'
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': dict(CONFIG, cpu_rows_per_step=n_rows, cpu_threads=cores),
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port-torch', 'sample': sample,
                             'tflops': cd_flops(n_rows) / sec / 1e12},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0,
            'note': 'TensorFlow 1.8 is not installable here; this is the torch-CPU restatement of the reference graph, pinned to '
                    'the numpy oracle, itself pinned by goldens generated from the reference sources (oracle/__init__.py)'}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as td
    from deepbelief_b200 import compat, ops, parallel
    from deepbelief_b200.layers.rbm import RBM
    from deepbelief_b200.layers.data import Data

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit('--gpus %d needs torchrun (one process per GPU)' % args.gpus)
    torch.cuda.set_device(local_rank)
    dev = 'cuda:%d' % local_rank
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        td.init_process_group(backend='nccl', device_id=torch.device(dev))
    n_local = N_GLOBAL // world
    K, Wm = args.steps, args.warmup
    peaks = load_peaks()

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    # ---- model: random-init weights of the named architecture, synthetic data of the named shape ----
    np.random.seed(0)
    torch.manual_seed(1234 + rank)
    rbm = RBM(V, H, W_std=0.01, name='bench_rbm', seed=7, device=dev)
    gen = torch.Generator(device=dev).manual_seed(0)                 # same weights on every rank
    rbm.set_weights(W=torch.randn(V, H, device=dev, generator=gen) * 0.01)
    optimizer = compat.AdamOptimizer(learning_rate=1e-3)
    pool = [(torch.rand(n_local, V, device=dev) < 0.5).float() for _ in range(2)]

    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    windows = []

    # ---- value: HBM-resident inputs ----
    rbm.rng.row_offset = rank * n_local
    opt_state = rbm.new_optimizer_state(optimizer)
    grad = torch.empty_like(rbm._params)
    chain_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]

    def one_step(i, timed_idx=None):
        if timed_idx is not None:
            chain_ev[timed_idx][0].record()
        # RBM.cd_update: one process = chain + dW with the optimizer and the split-weight refresh fused into its epilogue
        # (db_cd_tc_step_update); several = chain + dW, NCCL all-reduce, optimizer, refresh. The events bracket the whole update.
        rbm.cd_update(pool[i % 2], K_GIBBS, optimizer, opt_state, grad, n_global=N_GLOBAL)
        if timed_idx is not None:
            chain_ev[timed_idx][1].record()

    for i in range(Wm):
        one_step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_a = time.perf_counter()
    e0.record()
    for i in range(K):
        one_step(Wm + i, i)
    e1.record()
    barrier()
    t_b = time.perf_counter()
    windows.append((t_a, t_b))
    ms_total = parallel.max_over_ranks(e0.elapsed_time(e1), device=dev)
    chain_ms = sum(a.elapsed_time(b) for a, b in chain_ev) / K
    chain_ms = parallel.max_over_ranks(chain_ms, device=dev)
    ms_per_step = ms_total / K
    value = N_GLOBAL * K_GIBBS / (ms_per_step * 1e-3)
    counted = count_kernel_launches(lambda: one_step(0))               # one extra untimed step under the profiler
    if counted is not None:
        launches_per_step, launch_names = counted
        launches_how = 'counted: kernels of one step under torch.profiler (CUPTI), x steps'
    else:
        launches_per_step, launch_names = rbm.launches_per_cd_step(K_GIBBS) + 2, None
        launches_how = 'formula (profiler unavailable): chain + dW + bias/optimizer kernels'
    del opt_state

    # ---- e2e: through RBM.train, batches in pinned host memory ----
    host_rows = 2 * n_local
    host = torch.empty((host_rows, V), dtype=torch.float32).pin_memory()
    host.copy_(torch.cat(pool).cpu())
    data = Data.from_arrays(host.numpy(), batch_size=n_local, log_epochs=False, name='bench_data', pin=False)
    data._pinned_store = host
    data.shard(0, 1)               # every rank already holds ITS rows of the global batch: RBM.train must not shard them again
    monitor = []
    LAG = 3                                    # the host consumes step i's monitor value while steps i+1..i+3 are in flight
    mon_host = torch.zeros(LAG + 1, dtype=torch.float32).pin_memory()
    mon_ev = [torch.cuda.Event(blocking=True) for _ in range(LAG + 1)]     # blocking: a waiting rank yields its core
    mon_state = {'n': 0, 'read': 0}

    def step_callback(it, g):
        # per-step monitor, device->host EVERY step: L1 norm of the visible-bias gradient (|<v>_data - <v>_model|).
        # The 4-byte copy of step i is enqueued now and read on the host LAG steps later (a logging lag), so the host
        # keeps enqueueing while the device works and per-step host jitter (8 ranks share the box's cores) is absorbed.
        i = mon_state['n']
        s = i % (LAG + 1)
        mon_host[s:s + 1].copy_(g[V * H:V * H + V].abs().sum().reshape(1), non_blocking=True)
        mon_ev[s].record()
        mon_state['n'] = i + 1
        while mon_state['n'] - mon_state['read'] > LAG:
            r = mon_state['read'] % (LAG + 1)
            mon_ev[r].synchronize()
            monitor.append(float(mon_host[r]))
            mon_state['read'] += 1

    def drain_monitor():
        while mon_state['read'] < mon_state['n']:
            r = mon_state['read'] % (LAG + 1)
            mon_ev[r].synchronize()
            monitor.append(float(mon_host[r]))
            mon_state['read'] += 1

    def train(iters):
        rbm.checkpoint_iter = 1
        rbm.train(optimizer, data, None, num_gibbs_steps=K_GIBBS, pcd=False, max_iterations=iters,
                  step_callback=step_callback)

    train(Wm)
    drain_monitor()
    barrier()
    t_a = time.perf_counter()
    e0.record()
    train(K)
    drain_monitor()                                         # every step's monitor value reaches the host inside the window
    e1.record()
    barrier()
    t_b = time.perf_counter()
    windows.append((t_a, t_b))
    e2e_ms = parallel.max_over_ranks(e0.elapsed_time(e1), device=dev) / K
    e2e_value = N_GLOBAL * K_GIBBS / (e2e_ms * 1e-3)
    # the synthetic batch is {0,1}: the feeder stages it as bytes (Data.binary_bytes) and the chain's first kernel reads uint8
    h2d = rbm.last_feeder.h2d_bytes // K
    assert h2d in (n_local * V, n_local * V * 4) and rbm.last_feeder.h2d_bytes == h2d * K, \
        'feeder copied %d bytes over %d steps' % (rbm.last_feeder.h2d_bytes, K)

    # ---- second metric: GPLVM LML + gradient at N=5000 (replicas only across GPUs) ----
    del pool, data, host, grad
    rbm._tc_plan = None
    torch.cuda.empty_cache()
    gX = torch.randn(GP_N, GP_Q, device=dev)
    gY = (torch.rand(GP_N, GP_D, device=dev) < 0.13).float()
    gY -= gY.mean(dim=0, keepdim=True)
    isp = float(np.log(np.expm1(1.0)))
    hyp_opt = torch.tensor([isp, isp, isp], dtype=torch.float32, device=dev)
    gplan = ops.GplvmPlan(GP_N, GP_Q, GP_D, device=dev)
    gp_iters = max(3, min(K, 10))
    for _ in range(3):
        gplan.evaluate(gX, gY, hyp_opt, 7, jitter=1e-6)
    barrier()
    t_a = time.perf_counter()
    e0.record()
    for _ in range(gp_iters):
        gplan.evaluate(gX, gY, hyp_opt, 7, jitter=1e-6)
    e1.record()
    barrier()
    t_b = time.perf_counter()
    windows.append((t_a, t_b))
    gp_ms = parallel.max_over_ranks(e0.elapsed_time(e1), device=dev) / gp_iters
    gp_info = int(gplan.info.item())
    # the HBM-bound member of the chain, timed alone: RBF Gram (K3), 4 N^2 + 4 N Q algorithmic bytes per launch; the 100 MB
    # outputs cycle through the allocator's freed blocks (3 x 100 MB > L2)
    g_hyp = torch.tensor([1.0, 1.0, 1.0], dtype=torch.float32, device=dev)
    for _ in range(3):
        ops.gram_rbf(gX, None, g_hyp, add_noise=True)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(30):
        ops.gram_rbf(gX, None, g_hyp, add_noise=True)
    g1.record()
    torch.cuda.synchronize()
    gram_us = g0.elapsed_time(g1) / 30 * 1e3
    gram_gbs = (4.0 * GP_N * GP_N + 4.0 * GP_N * GP_Q) / 1e9 / (gram_us * 1e-6)
    gp_flop = float(GP_N) ** 3 + 3.0 * GP_N * GP_N * GP_D
    gram_traffic = ncu_dram_bytes('r01_gram_full_raw.csv', 'gram_kernel')

    # GPLVM end to end through the layer API (gplvm.py:414-468): Y and X0 come from HOST memory into the constructor, the
    # optimisation loop runs `gp_e2e_iters` iterations (LML+grad + Adam on {X, hyper-parameters}), the final loss is read back
    del gplan
    torch.cuda.empty_cache()
    from deepbelief_b200.layers.gplvm import GPLVM
    Y_host = (gY + 0.0).cpu().numpy()
    X_host = gX.cpu().numpy()
    gp_e2e_iters = 100          # 1 % of the reference example's run (examples/gplvm_horses/flags.py:35, num_iterations = 10000)

    def gplvm_e2e():
        m = GPLVM(Y_host, GP_Q, X0=X_host, name='bench_gplvm', device=dev)
        m.optimize(compat.AdamOptimizer(learning_rate=1e-3), num_iterations=gp_e2e_iters)
        return float(m.loss_and_grad()[0].item())

    gplvm_e2e()                                                       # warm-up (allocator, graph capture path)
    gplvm_e2e()                                                       # (the second construction in a process is still ~80 ms slower)
    # three whole runs, the median reported: a run builds and drops a 1.3 GB model, and one cyclic-garbage collection or
    # allocator round trip landing inside a run (observed once: +200 ms) is not a property of the path
    import gc
    gp_e2e_runs = []
    for _ in range(3):
        gc.collect()
        barrier()
        t_a = time.perf_counter()
        e0.record()
        gp_e2e_loss = gplvm_e2e()
        e1.record()
        barrier()
        t_b = time.perf_counter()
        windows.append((t_a, t_b))
        gp_e2e_runs.append(parallel.max_over_ranks(e0.elapsed_time(e1), device=dev) / (gp_e2e_iters + 1))
    gp_e2e_ms = sorted(gp_e2e_runs)[1]

    # ---- GPDBN (BASELINE configs[3]): synthetic binary 5000x784, stack 784->200->100->GPRBM(50), Q=2, T=0.1 ----
    gpdbn = None
    if world == 1:
        del gX, gY
        torch.cuda.empty_cache()
        from deepbelief_b200.layers.gprbm import GPRBM
        from deepbelief_b200.layers.gplvm import SEKernel
        rs = np.random.RandomState(0)
        Vd = (rs.rand(GP_N, GP_D) < 0.13).astype(np.float32)
        X0 = rs.normal(size=(GP_N, GP_Q)).astype(np.float32)
        l0 = RBM(GP_D, 200, W_std=0.05, temperature=0.1, name='gpdbn_l0', seed=11, device=dev)
        l1 = RBM(200, 100, W_std=0.05, temperature=0.1, bottom=l0, name='gpdbn_l1', seed=12, device=dev)
        top = GPRBM(100, 50, GP_Q, GP_N, False, X0=X0, V=Vd, kern=SEKernel(alpha=False, gamma=False, device=dev),
                    noise_variance=0.01, fixed_noise_variance=True, W_std=0.05, temperature=0.1, bottom=l1,
                    name='gpdbn_top', seed=13, device=dev)
        adam = compat.AdamOptimizer(learning_rate=1e-3)
        top.optimize(adam, num_iterations=2)
        torch.cuda.synchronize()
        n_it = 5
        t_a = time.perf_counter()
        e0.record()
        top.optimize(adam, num_iterations=n_it)
        e1.record()
        torch.cuda.synchronize()
        t_b = time.perf_counter()
        windows.append((t_a, t_b))
        gpdbn = {'metric': 'GPDBN train iterations/s (loss + full-stack gradient + Adam)', 'value': n_it * 1e3 / e0.elapsed_time(e1),
                 'unit': 'iterations/s', 'ms_per_iteration': e0.elapsed_time(e1) / n_it, 'N': GP_N,
                 'stack': '784-200-100-GPRBM(50), Q=2, T=0.1, noise 0.01 fixed', 'data': 'synthetic Bernoulli(0.13)'}

    clocks = None
    if sampler is not None:
        time.sleep(0.15)
        sampler.stop()
        clocks = sampler.summary(windows)

    # ---- CPU baselines (rank 0, N=1 only): bounded samples of the same steps, torch-CPU ops on all host threads ----
    cpu = None
    gp_cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sec, cores = cpu_cd_steps_torch(CPU_ROWS, steps=3, warmup=1)
        sec_np, _ = cpu_cd_steps_numpy(CPU_ROWS, steps=1, warmup=1)
        cpu = {'value': CPU_ROWS * K_GIBBS / sec, 'unit': UNIT, 'cores': cores, 'kind': 'port-torch',
               'sample': cd_cpu_sample_text(CPU_ROWS) + ', 3 timed steps after 1 warm-up; torch-CPU fp32 (MKL sgemm) restatement, '
                         'oracle/torch_port.py',
               'tflops': cd_flops(CPU_ROWS) / sec / 1e12,
               'numpy_port': {'value': CPU_ROWS * K_GIBBS / sec_np, 'unit': UNIT, 'kind': 'port',
                              'sample': 'same step with the numpy oracle (oracle/rbm.py), 1 timed step'}}
        gsec, gcores = cpu_gplvm_evals(evals=2, warmup=1)
        gp_cpu = {'value': 1.0 / gsec, 'unit': 'evals/s', 'cores': gcores, 'kind': 'port-torch', 'ms_per_eval': gsec * 1e3,
                  'sample': 'the full N=5000, D=784, Q=2 evaluation (not a sub-sample), 2 timed evaluations after 1 warm-up: '
                            'torch-CPU fp32 Gram + linalg.cholesky + solve_triangular + autograd (oracle/torch_port.py)'}

    # dram bytes per launch of the dominant kernel from the committed round-2 ncu capture: mean over the captured launches
    # (a K = 8192 `down`, a K = 4096 `up`, the fused dW launch)
    traffic = None
    tr = [ncu_dram_bytes('r02_prof_tc_gemm_raw.csv', 'tc_gemm_kernel', nth=i) for i in range(3)]
    tr = [t for t in tr if t is not None]
    if tr:
        traffic = sum(tr) / len(tr)
    if rank == 0:
        achieved = cd_flops(n_local) / (chain_ms * 1e-3) / 1e12
        peak = peaks['tensor_sustained']
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': Wm,
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'f32 (fp16 hi/lo split operands, fp32 TMEM accumulate)', 'data': 'synthetic',
            'config': dict(CONFIG, parallelism=('dp1 (one process: optimizer fused into the dW kernel)' if world == 1 else
                                                'dp%d (rows sharded; %s)' % (world, 'dW reduce-scattered by peer stores from the dW epilogue over NVLink, '
                                                'sharded update, rows pushed back; NCCL all-reduce of db|dc' if getattr(rbm, '_peer', None)
                                                else 'NCCL sum-all-reduce of dW|db|dc')),
                           local_batch=n_local),
            'e2e': {'value': e2e_value, 'unit': UNIT, 'ms_per_step': e2e_ms, 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': 4, 'h2d_dtype': 'uint8' if h2d == n_local * V else 'float32', 'api': 'RBM.train(AdamOptimizer, Data.from_arrays(pinned host), '
                    'num_gibbs_steps=10, pcd=False)', 'monitor_last': monitor[-1] if monitor else None},
            'gpu_launches': launches_per_step * K,
            'gpu_launches_how': launches_how, 'gpu_launch_kernels': launch_names,
            'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
                         'frac': achieved / peak, 'traffic': traffic,
                         'traffic_unit': 'bytes per tc_gemm_kernel launch (ncu dram read+write, mean of the three launches in '
                                         'profiles/r02_prof_tc_gemm_raw.csv; algorithmic bytes of a chain launch ~0.54 GB)',
                         'kernel': 'tc_gemm_kernel (22 launches per step: 21 chain GEMMs + dW with the fused update)',
                         'peak_source': 'MEASURED_PEAKS.json bf16 sustained (%s); fp16 runs at the bf16 rate'
                                        % peaks['source'],
                         'executed_tflops': 2.0 * achieved,
                         # north_star names the TF32 tensor peak as the roofline of the CD GEMMs (a single-pass fp32-accurate
                         # product would run as kind::tf32, half the 16-bit rate): the same achieved figure against it
                         'frac_of_tf32_peak': achieved / (0.5 * peak),
                         'frac_executed_of_peak': 2.0 * achieved / peak,
                         'note': 'achieved counts ALGORITHMIC flop (2k+3)*2*N*V*H; every MMA is issued twice '
                                 '(hi and lo half of the split real operand), so the tensor pipe executes 2x',
                         'chain_ms_per_step': chain_ms,
                         'timed_region': 'db_cd_tc_step_update call (chain + dW with fused optimizer and cache rewrite)' if world == 1 else
                                         'whole update incl. NCCL all-reduce and optimizer (conservative)'},
            'cpu_baseline': cpu,
            'clocks': clocks,
            'gplvm': {'metric': 'GPLVM LML+grad evals/s at N=5000', 'value': world * 1e3 / gp_ms, 'unit': 'evals/s',
                      'ms_per_eval': gp_ms, 'N': GP_N, 'D': GP_D, 'Q': GP_Q, 'info': gp_info, 'scaling': 'replicas only',
                      'algorithmic_tflops': gp_flop / (gp_ms * 1e-3) / 1e12,
                      # N^3 + 3 N^2 D algorithmic FLOP against the TF32 tensor peak north_star names = half the
                      # measured sustained bf16 rate (kind::tf32 runs at half the 16-bit rate); the N^3-class stages run as 3xTF32
                      # tcgen05 GEMMs, i.e. the tensor pipe executes ~3x the algorithmic product FLOP
                      'roofline': {'bound': 'tensor', 'achieved': gp_flop / (gp_ms * 1e-3) / 1e12,
                                   'peak': 0.5 * peaks['tensor_sustained'], 'unit': 'TFLOP/s',
                                   'frac': gp_flop / (gp_ms * 1e-3) / 1e12 / (0.5 * peaks['tensor_sustained']),
                                   'peak_source': '0.5 x MEASURED_PEAKS.json bf16 sustained (%s) = TF32 dense' % peaks['source'],
                                   'frac_of_fp32_ffma_peak': gp_flop / (gp_ms * 1e-3) / (148 * 128 * 2 * 1.965e9),
                                   'note': 'whole evaluation (Gram, Cholesky, triangular inverse, K^-1, K^-1 Y, A A^T, contraction), '
                                           'not one kernel; see profiles/ for the per-kernel shares'},
                      'e2e': {'value': world * 1e3 / gp_e2e_ms, 'unit': 'evals/s', 'ms_per_eval': gp_e2e_ms,
                              'api': 'GPLVM(Y_host, Q, X0=X_host) + GPLVM.optimize(AdamOptimizer, num_iterations=%d) + loss read-back; '
                                     'the constructor (host->device copy of Y and X0, workspace allocation) is inside the timed region and '
                                     'amortised over the %d evaluations; median of 3 such runs' % (gp_e2e_iters, gp_e2e_iters + 1),
                              'ms_per_eval_runs': gp_e2e_runs,
                              'h2d_bytes_per_eval': (GP_N * GP_D + GP_N * GP_Q) * 4 / (gp_e2e_iters + 1), 'd2h_bytes_per_eval': 4 / (gp_e2e_iters + 1),
                              'final_loss': gp_e2e_loss},
                      'cpu_baseline': gp_cpu,
                      'gram': {'kernel': 'gram_kernel<2>', 'us_per_launch': gram_us,
                               'roofline': {'bound': 'hbm', 'achieved': gram_gbs, 'peak': load_peaks()['hbm_gbs'], 'unit': 'GB/s',
                                            'frac': gram_gbs / load_peaks()['hbm_gbs'],
                                            'traffic': gram_traffic,
                                            'traffic_note': 'dram read+write inside the kernel, read from profiles/r01_gram_full_raw.csv; the '
                                                            '126 MB L2 absorbs the rest of the 100 MB it writes'}}},
            'gpdbn': gpdbn,
        }
        print(json.dumps(line))
    if world > 1:
        td.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg (profiling runs)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == 'reference':
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = args.warmup          # the contract asks for W >= 3; the driver passes it explicitly
    return run_b200(args)


if __name__ == '__main__':
    sys.exit(main())
