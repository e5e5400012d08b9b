"""Benchmark of the hot path on the workload BASELINE.json quotes its scaling metric on (configs[2]):
fixed-K preconditioned-CG solves with GGN-vector products on Martens' deep autoencoder
784-1000-500-250-30-250-500-1000-784 (sigmoid units, linear code layer, BCE-with-logits, mean reduction),
batch 60 000 as 8 `acc_step` chunks of 7 500, chunk c on rank c mod N.

    python bench.py --gpus 1 --steps 20 --warmup 3            # this repo's sm_100a path
    python bench.py --impl reference --steps 5 --warmup 1     # the CPU oracle port of the reference path
    torchrun --nproc-per-node N ... bench.py --gpus N ...     # the same 60 000 samples sharded over N ranks (strong scaling)

One "step" = one PCG solve of (G + lambda I) x = -g with exactly K_cg = 50 iterations (tol = 0, Martens' criterion
off, so only the iteration cap stops it -- BASELINE.md section 3; all termination quantities are still computed every
iteration).  Each iteration = one GGN-vector product over the WHOLE 60 000-sample batch (the rank's chunks, then one
all-reduce of the P-vector when N > 1) + one fused CG vector update.  The unit counted is that full-batch product, so
`value` = steps * K_cg / seconds at every N ("scaling": "strong").

Prints ONE JSON line (rank 0).  Timing: CUDA events on the launching stream around every step, max over ranks, L2
flushed (256 MiB write) between steps outside the timed region.  Extra blocks: `e2e` (host buffers in, Newton step
out), `e2e_api` (HessianFree.get_preconditioner + acc_step on host tensors), `roofline` (dominant contraction timed
alone), `roofline_cg_update`, `allreduce`, `cfg2` (the round-1 workload, BASELINE.json configs[1], N = 1 only),
`cpu_baseline`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

AE_WIDTHS = [784, 1000, 500, 250, 30, 250, 500, 1000, 784]
AE_LINEAR_AFTER = 3          # the 30-unit code layer is linear (Martens 2010)
BATCH, CHUNKS = 60000, 8     # 7 500 samples per chunk = the per-GPU shard at 8 GPUs
MLP_WIDTHS, MLP_BATCH = [784, 512, 512, 10], 4096   # BASELINE.json configs[1] (extra block)
K_CG = 50
DAMPING = 1e-3  # small enough that 50 iterations stay clear of the float32 rounding floor (with lambda = 1 the
#                 solve converges in ~13 iterations and a fixed-50 run would divide 0/0)
METRIC = "GGN-vector products/sec (CG iters/sec)"
UNIT = "products/s"
WORKLOAD = "martens_ae_784-1000-500-250-30-250-500-1000-784_sigmoid_bce_batch60000_ggn_fisher_pcg"


def build_net(widths, act, linear_after=(), seed=0):
    torch.manual_seed(seed)
    mods = []
    for i in range(len(widths) - 1):
        mods.append(torch.nn.Linear(widths[i], widths[i + 1]))
        if i < len(widths) - 2 and i not in linear_after:
            mods.append(act())
    return torch.nn.Sequential(*mods)


def build_ae(seed=0):
    return build_net(AE_WIDTHS, torch.nn.Sigmoid, (AE_LINEAR_AFTER,), seed)


def build_mlp(seed=0):
    return build_net(MLP_WIDTHS, torch.nn.ReLU, (), seed)


def ae_chunk(c, rows=BATCH // CHUNKS):
    """Chunk c of the synthetic batch: U[0,1) inputs, which double as the targets."""
    g = torch.Generator().manual_seed(1000 + c)
    return torch.rand(rows, AE_WIDTHS[0], generator=g)


def mlp_data(seed=1, n=MLP_BATCH):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n, MLP_WIDTHS[0], generator=g), torch.randint(0, MLP_WIDTHS[-1], (n,), generator=g)


def flops_per_product(widths, n):
    """F_Gv = 2N(4S - 2 m1): contractions only."""
    m = [widths[i] * widths[i + 1] for i in range(len(widths) - 1)]
    return 2.0 * n * (4 * sum(m) - 2 * m[0])


def config(n_gpus, params):
    """Identical in both arms (the driver compares it)."""
    return dict(workload=WORKLOAD, batch=BATCH, chunks=CHUNKS, params=params, cg_iters_per_step=K_CG, damping=DAMPING,
                l2="flushed between steps (256 MiB write)",
                parallelism=f"dp{n_gpus}: the 8 chunks dealt round-robin to the ranks, one all-reduce of the P-vector per CG iteration")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained"), src="measured")
    return dict(hbm=6650.0, tf=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """SM clock and clock-event (throttle) reasons of this rank's board while the timed region runs, every 50 ms.
    Sampled by a thread of THIS process through NVML (the library nvidia-smi prints from), initialised before the
    warm-up: a `nvidia-smi -lms` child per rank needs about a second to start on an 8-GPU box and holds driver locks
    while it does, which inside a 0.3 s timed region slowed the very thing it was watching (profiles/r2_summary.md).
    Falls back to that child, started at construction so that its start-up is over before the timed region, when the
    NVML binding is missing.  Only samples taken between __enter__ and __exit__ count."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.rows, self.proc, self.nvml, self.handle = [], None, None, None
        self.live, self.stop, self.thread = False, False, None
        try:
            import pynvml

            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
            except Exception:  # noqa: BLE001 -- old torch / MIG naming: NVML order == CUDA order without a device mask
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.source = "nvml"
            self.thread = threading.Thread(target=self._poll, daemon=True)
        except Exception:  # noqa: BLE001
            self.nvml = None
            self.source = "nvidia-smi"
            try:
                self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                              "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.thread = threading.Thread(target=self._pump, daemon=True)
            except OSError:
                self.proc = None
        if self.thread:
            self.thread.start()

    def _poll(self):
        n = self.nvml
        while not self.stop:
            if self.live:
                try:
                    mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                    mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                    self.rows.append((mhz, self.max_mhz, [name for name, bit in self.BITS if mask & bit]))
                except Exception:  # noqa: BLE001
                    pass
            time.sleep(0.05 if self.live else 0.005)

    def _pump(self):
        for line in self.proc.stdout:
            if not self.live:
                continue
            c = [x.strip() for x in line.split(",")]
            try:
                self.rows.append((float(c[0]), float(c[1]),
                                  [name for (name, _), v in zip(self.BITS, c[3:7]) if v.lower().startswith("active")]))
            except (ValueError, IndexError):
                continue

    def __enter__(self):
        self.live = True
        return self

    def __exit__(self, *a):
        self.live, self.stop = False, True
        if self.proc:
            self.proc.terminate()
        if self.thread:
            self.thread.join(timeout=2)

    def summary(self):
        sm = sorted(r[0] for r in self.rows)
        mx = max((r[1] for r in self.rows), default=0.0)
        reasons = sorted({name for r in self.rows for name in r[2]})
        busy = [v for v in sm if v > 0.5 * mx] or sm
        return dict(sm_mhz=busy[len(busy) // 2] if busy else None, sm_max_mhz=mx or None, reasons=reasons, samples=len(sm),
                    source=self.source)


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path (reference cg.py + _Gv through the BackPACK recipe)
# ------------------------------------------------------------------------------------------------------
CPU_SAMPLE_ROWS = 1500  # bounded sample: 1/40 of the batch, same net, same 50 iterations per solve


def cpu_solve_rate(solves, rows=CPU_SAMPLE_ROWS):
    """Full-batch-equivalent products/s of the reference path on this host: K_CG-iteration solves on `rows` samples of
    the workload; the cost of the reference's chunk loop is linear in the rows (optimizer.py:658-684), so the rate is
    scaled by rows / BATCH."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import hf_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)  # all host threads, whatever OMP_NUM_THREADS the launcher exported
    model = build_ae(0)
    loss_fn = torch.nn.BCEWithLogitsLoss()
    x = ae_chunk(0, rows)
    params = list(model.parameters())
    out = model(x)
    loss = loss_fn(out, x)
    grad = O.flatten(torch.autograd.grad(loss, params, create_graph=True)).detach()
    M = O.diag_precond(torch.rand_like(grad) * 1e-3, DAMPING)  # same kind of operator; values do not affect cost
    A = lambda v: O.Gv(loss, out, params, v) + DAMPING * v  # noqa: E731
    O.pcg(A, -grad, M=M, max_iter=2, tol=0.0)  # warm-up
    t0 = time.perf_counter()
    for _ in range(solves):
        O.pcg(A, -grad, M=M, max_iter=K_CG, tol=0.0)
    dt = time.perf_counter() - t0
    # the reference's cg spends K+1 products for K iterations (cg.py:188)
    return solves * K_CG * (rows / BATCH) / dt, dt, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    params = sum(p.numel() for p in build_ae(0).parameters())
    cpu_solve_rate(max(1, min(args.warmup, 2)))
    rate, dt, threads = cpu_solve_rate(args.steps)
    line = dict(metric=METRIC, value=rate, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * dt / args.steps / (CPU_SAMPLE_ROWS / BATCH), higher_is_better=True, scaling="strong",
                vs_baseline=None, dtype="f32", data="synthetic", impl="reference", config=config(args.gpus, params),
                cpu_baseline=dict(value=rate, unit=UNIT, cores=threads, kind="port",
                                  sample=f"{args.steps} solves x {K_CG} CG iterations on {CPU_SAMPLE_ROWS} of the {BATCH} samples "
                                         f"({dt:.1f} s), rate scaled by {CPU_SAMPLE_ROWS}/{BATCH}; oracle port of reference cg.py + _Gv "
                                         "(BackPACK autograd recipe)"),
                e2e=dict(value=rate, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def run_native(args):
    import torch.distributed as dist

    from pytorchhessianfree_b200 import DiagonalPreconditioner, HessianFree, _lib, pcg_device
    from pytorchhessianfree_b200.lowering import lower_module
    from pytorchhessianfree_b200.native import NativeNet
    from pytorchhessianfree_b200.problem import NativeProblem

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    lib = _lib.load()
    engine = args.engine

    model = build_ae(0).to(dev)
    loss_fn = torch.nn.BCEWithLogitsLoss()
    params = list(model.parameters())
    prog = lower_module(model, loss_fn, params)
    theta = torch.cat([p.detach().reshape(-1) for p in params]).contiguous()
    P = theta.numel()
    net = NativeNet(prog.layers, prog.loss, prog.reduction, prog.n_params, engine=engine)
    mine = [c for c in range(CHUNKS) if c % world == rank]
    host = [ae_chunk(c).pin_memory() for c in mine]
    resident = [h.to(dev) for h in host]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def setup(chunks, grp):
        prob = NativeProblem(net, theta, "ggn", [(x, x) for x in chunks], group=grp)
        prob.linearize()
        g = prob.gradient()
        M = DiagonalPreconditioner(prob.fisher_diag(), DAMPING)
        return prob, g, M

    def solve(prob, g, M):
        return pcg_device(prob.matvec, -g, minv=M.minv, damping=DAMPING, max_iter=K_CG, tol=0.0,
                          martens_conv_crit=False, store_x_at_iters=None, poll=K_CG, out_buffer=prob.out_buffer())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm: inputs in HBM, linearisation done, time the solves -------------------
    clk = ClockSampler(local)  # initialised here, armed around the timed region
    prob, g, M = setup(resident, group)
    for _ in range(args.warmup):
        solve(prob, g, M)
    barrier()
    n0 = lib.hf_debug_launch_count()
    times = []
    with clk:
        for _ in range(args.steps):
            flush.fill_(1)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            xs, _, why = solve(prob, g, M)
            e1.record()
            barrier()
            times.append(e0.elapsed_time(e1))
        launches = lib.hf_debug_launch_count() - n0
    assert why == "Number of iterations" and len(xs) == K_CG + 1, f"fixed-K solve stopped early: {why}, {len(xs) - 1} iterations"
    total_ms = reduce_max(sum(times))
    value = args.steps * K_CG / (total_ms * 1e-3)
    x_final = xs[-1].clone()

    # ---- data-parallel correctness on the hardware: replicas bit-identical, and equal to the unsharded solve ----
    checks = None
    if world > 1:
        lo, hi = x_final.clone(), x_final.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        identical = bool(torch.equal(lo, hi))
        # every rank repeats the solve on all 8 chunks by itself (no collective) and compares
        full = [ae_chunk(c).to(dev) for c in range(CHUNKS)]
        p1, g1, M1 = setup(full, None)
        xs1 = solve(p1, g1, M1)[0]
        rel = lambda u, w: float(((u.double() - w.double()).norm() / w.double().norm()).item())  # noqa: E731
        err10, err = rel(xs[10], xs1[10]), rel(x_final, xs1[-1])
        del p1, full
        assert identical, "data-parallel replicas diverged: final CG iterates differ between ranks"
        # CG iterates within rtol 1e-3 (north_star) where that is meaningful: iterate 10.  The 50th iterate of this
        # ill-conditioned solve (lambda = 1e-3) moves by ~1e-3 under ANY change of summation order (SURVEY.md section 7,
        # hard part 4), so it is reported and only bounded loosely.
        assert err10 < 1e-3, f"sharded solve differs from the single-GPU solve at iteration 10: rel L2 {err10:.2e}"
        assert err < 1e-2, f"sharded solve differs from the single-GPU solve at iteration {K_CG}: rel L2 {err:.2e}"
        checks = dict(replicas_bit_identical=identical, rel_l2_vs_unsharded_iter10=err10, rel_l2_vs_unsharded_iter50=err)

    # ---- end-to-end arm: host buffers in, result out, everything inside the timed region -----------
    def e2e_step():
        chunks = [h.to(dev, non_blocking=True) for h in host]
        pr, gg, MM = setup(chunks, group)
        xs_, _, _ = solve(pr, gg, MM)
        return xs_[-1].cpu()  # the Newton step (P floats) read back

    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    barrier()
    e2e_steps = max(3, args.steps // 4)
    e2e_ms = 0.0
    for _ in range(e2e_steps):
        flush.fill_(1)
        barrier()
        t0 = time.perf_counter()
        e2e_step()
        torch.cuda.synchronize()
        e2e_ms += 1e3 * (time.perf_counter() - t0)
    e2e_value = e2e_steps * K_CG / (reduce_max(e2e_ms) * 1e-3)

    # ---- the same through the optimizer's public API: get_preconditioner + acc_step on host tensors ----
    def api_steps(mdl, lfn, datalist, first, n_steps, grp):
        import warnings

        opt = HessianFree(mdl.parameters(), process_group=grp)
        ms, iters = [], []
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for i in range(n_steps + 1):
                barrier()
                t0 = time.perf_counter()
                Mf = opt.get_preconditioner(mdl, lfn, first[0], first[1], "mean")
                opt.acc_step(mdl, lfn, datalist, M_func=Mf)
                torch.cuda.synchronize()
                if i:  # the first step pays lazy initialisation
                    ms.append(1e3 * (time.perf_counter() - t0))
                    iters.append(opt.state["num_cg_iters"][-1])
        t = reduce_max(sum(ms))
        return dict(ms_per_step=t / n_steps, cg_iters=iters, value=sum(iters) / (t * 1e-3), unit=UNIT,
                    what="HessianFree.get_preconditioner (first local chunk) + acc_step on pinned host chunks: H2D, linearise, "
                         "gradient, Fisher, PCG with Martens' criterion, cg-backtracking, line search, parameter update")

    api_model = build_ae(0).to(dev)
    e2e_api = api_steps(api_model, loss_fn, [(h, h) for h in host], (host[0], host[0]), 3, group)

    # ---- the one collective of the path, alone ----
    ar = None
    if world > 1:
        buf = torch.randn(P, device=dev)
        from pytorchhessianfree_b200.dist import all_reduce_sum
        for _ in range(5):
            all_reduce_sum(buf, group)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            all_reduce_sum(buf, group)
        e1.record()
        barrier()
        us = reduce_max(e0.elapsed_time(e1)) * 1e3 / 50
        ar = dict(bytes=4 * P, nccl_us=us, nccl_bus_gbs=2.0 * (world - 1) / world * 4 * P / (us * 1e-6) / 1e9,
                  note="all-reduce of the FP32 P-vector on the launching stream, back to back: ncclAllReduce, and the "
                       "switch-reduced kernel of this repository (hf_allreduce_multimem) the solve uses when the fabric has "
                       "multicast; reference: 725 GB/s bus at 1 GiB")
        from pytorchhessianfree_b200.dist import SymmetricVector
        sv = SymmetricVector.try_create(P, dev, group, force=True)
        ar["solve_uses"] = "multimem" if prob.out_buffer() is not None else "nccl"
        if sv is not None:
            sv.vec.copy_(buf)
            for _ in range(5):
                sv.all_reduce_()
            barrier()
            e0.record()
            for _ in range(50):
                sv.all_reduce_()
            e1.record()
            barrier()
            us = reduce_max(e0.elapsed_time(e1)) * 1e3 / 50
            ar.update(us=us, bus_gbs=2.0 * (world - 1) / world * 4 * P / (us * 1e-6) / 1e9)
            sv.vec.zero_()
        else:
            ar.update(us=ar["nccl_us"], bus_gbs=ar["nccl_bus_gbs"])

    # ---- per-kernel rooflines and the round-1 workload, measured live (rank 0, kernels alone on the device) ----
    pk = peaks()
    roof, extra = None, {}
    if rank == 0:
        # rank-local problem: the per-kernel timings must not issue collectives the other ranks do not join
        local_prob = NativeProblem(net, theta, "ggn", [(resident[0], resident[0])], group=None)
        roof, extra = kernel_rooflines(lib, local_prob, theta, dev, pk, engine)
        if world == 1:
            extra["cfg2"] = cfg2_block(lib, dev, pk, engine, api_steps, barrier)
            del prob, local_prob, resident
            torch.cuda.empty_cache()
            extra["conv"] = conv_block(dev, engine)
    if world > 1:
        dist.barrier()

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            rate, dt, threads = cpu_solve_rate(2)
            cpu = dict(value=rate, unit=UNIT, cores=threads, kind="port",
                       sample=f"2 solves x {K_CG} CG iterations on {CPU_SAMPLE_ROWS} of the {BATCH} samples ({dt:.1f} s), rate scaled "
                              f"by {CPU_SAMPLE_ROWS}/{BATCH}")
        cfg = config(world, P)
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=total_ms / args.steps, higher_is_better=True, scaling="strong", vs_baseline=None,
                    dtype="f32", data="synthetic", config=cfg, engine=engine,
                    clocks=clk.summary(), gpu_launches=int(launches),
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=int(sum(h.numel() for h in host) * 4),
                             d2h_bytes_per_step=int(P * 4),
                             what="H2D of the rank's chunks + forward + gradient + Fisher diagonal + 50-iteration PCG + D2H step"),
                    e2e_api=e2e_api, roofline=roof, cpu_baseline=cpu, allreduce=ar, dp_checks=checks, **extra)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _timed(fn, flush, reps=10):
    """Median CUDA-event time (ms) of fn alone on the device, L2 flushed before every launch."""
    fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(400_000)  # ~0.2 ms of device spin: the host enqueues e0/fn/e1 behind it, so the
        e0.record()                 # interval holds device time only, not Python launch latency
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[len(ms) // 2]


def contraction_roofline(lib, dev, pk, engine, M, N, K, what, flush):
    """One tensor-tile contraction C[M,N] = A[M,K] B[N,K]^T through the C entry point the products use, timed alone.
    On the tensor engine the pair kernel reads pre-split operand images, built once (engine 2) and then reused
    (engine 3), exactly as a solve reuses the images of its linearisation."""
    from pytorchhessianfree_b200._lib import Operand
    a = torch.randn(M, K, device=dev)
    b = torch.randn(N, K, device=dev)
    c = torch.empty(M, N, device=dev)
    A = (Operand * 1)(Operand(a.data_ptr(), K, 1))
    B = (Operand * 1)(Operand(b.data_ptr(), K, 1))
    stream = torch.cuda.current_stream().cuda_stream
    kernel = "gemm_simt_kernel"
    eng, ws_ptr, ws_bytes = 0, None, 0
    if engine == "tc":
        ws_bytes = lib.hf_contract_workspace_bytes(M, N, K, 1)
        ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=dev)
        ws_ptr = (ws.data_ptr() + 255) // 256 * 256
        pair = os.environ.get("HF_TC2", "1") != "0" and M * N >= 74 * 256 * 256 * 0.75
        if pair and lib.hf_contract(2, M, N, K, 1, A, B, c.data_ptr(), N, ws_ptr, ws_bytes, stream) == 0:
            eng, kernel = 3, "gemm_tc2_kernel (256x256 CTA-pair tiles, pre-split operand images)"
        else:
            eng, kernel = 1, "gemm_tc_kernel (128x128 tiles, in-kernel splitter)"

    def run():
        rc = lib.hf_contract(eng, M, N, K, 1, A, B, c.data_ptr(), N, ws_ptr, ws_bytes, stream)
        assert rc == 0, lib.hf_last_error_string()
    t = _timed(run, flush)
    f = 2.0 * M * N * K
    tf = f / (t * 1e-3) / 1e12
    return dict(bound="tensor", achieved=tf, peak=pk["tf"], unit="TFLOP/s", frac=tf / pk["tf"],
                # dram__bytes_read.sum + dram__bytes_write.sum of exactly this launch (L2 flushed before it) from one
                # `ncu --set full` capture (tools/roofline_kernel.py, profiles/r2_summary.md): 62.1 MB read + 8.3 MB written
                # (most of the 30 MB of C is still in the L2 when the kernel ends) against 83.5 MB of operand forms + C
                traffic=70.45e6 if (eng == 3 and (M, N, K) == (7500, 1000, 784)) else None,
                traffic_source="ncu --set full, gpurun_out/prof_r2_roofline.ncu-rep, summarised in profiles/r2_summary.md",
                peak_source=pk["src"], kernel=f"{kernel}: contraction {M}x{N}x{K} ({what})", us_per_launch=1e3 * t,
                flops_per_launch=f,
                note="split precision (TF32 main term + two BF16 correction terms = 4 bf16-MMA equivalents per product), "
                     "so the algorithmic ceiling is 1/4 of the bf16 peak")


def kernel_rooflines(lib, prob, theta, dev, pk, engine):
    """Time (a) one GGN-vector product on one 7 500-sample chunk and its dominant contraction, (b) the fused CG vector
    update, each alone on the device with CUDA events, L2 flushed between launches."""
    from pytorchhessianfree_b200._lib import PCG_FUSED
    from pytorchhessianfree_b200.cg import _Solver

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    prob.linearize()
    v, out = torch.randn_like(theta), torch.empty_like(theta)
    n = prob.mvp_lins[0].n
    t_mv = _timed(lambda: prob.matvec(v, out), flush)
    f_mv = flops_per_product(AE_WIDTHS, n)
    # dominant kernel: the layer-1 R-op forward  Rz1[7500,1000] = a0[7500,784] V1[1000,784]^T  heads the launch list of
    # a product (profiles/r2_launches_cfg3.csv)
    roof = contraction_roofline(lib, dev, pk, engine, n, AE_WIDTHS[1], AE_WIDTHS[0], "layer-1 R-op forward of one chunk", flush)
    upd = {}
    for name, P in (("cfg3_P2837314", theta.numel()), ("cfg2_P669706", 669706)):
        b = torch.randn(P, device=dev)
        s = _Solver(b, 10 ** 6)
        minv = torch.rand(P, device=dev) + 0.5
        Bp = torch.randn(P, device=dev)
        s.init(None, None, minv, DAMPING, 0.0, None, False, False)
        t_u = _timed(lambda: s.iterate(PCG_FUSED, Bp=Bp, minv=minv, lam=DAMPING), flush, reps=20)
        gbs = 36.0 * P / (t_u * 1e-3) / 1e9
        upd[name] = dict(bound="hbm", achieved=gbs, peak=pk["hbm"], unit="GB/s", frac=gbs / pk["hbm"],
                         us_per_launch=1e3 * t_u, bytes_per_launch=36 * P)
    extra = dict(roofline_cg_update=upd,
                 matvec=dict(us_per_product=1e3 * t_mv, rows=n, tflops_algorithmic=f_mv / (t_mv * 1e-3) / 1e12,
                             flops_per_product=f_mv, frac_of_bf16_peak=f_mv / (t_mv * 1e-3) / 1e12 / pk["tf"],
                             what="one GGN-vector product on one 7 500-sample chunk (1/8 of a full-batch product)"))
    return roof, extra


def cfg2_block(lib, dev, pk, engine, api_steps, barrier):
    """BASELINE.json configs[1], the round-1 bench workload, for continuity: MLP 784-512-512-10 ReLU, CE, batch 4096."""
    from pytorchhessianfree_b200 import DiagonalPreconditioner, pcg_device
    from pytorchhessianfree_b200.lowering import lower_module
    from pytorchhessianfree_b200.native import NativeNet
    from pytorchhessianfree_b200.problem import NativeProblem

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    model = build_mlp(0).to(dev)
    loss_fn = torch.nn.CrossEntropyLoss()
    params = list(model.parameters())
    prog = lower_module(model, loss_fn, params)
    theta = torch.cat([p.detach().reshape(-1) for p in params]).contiguous()
    net = NativeNet(prog.layers, prog.loss, prog.reduction, prog.n_params, engine=engine)
    xh, th = mlp_data()
    x, t = xh.to(dev), th.to(dev)
    prob = NativeProblem(net, theta, "ggn", [(x, t)])
    prob.linearize()
    g = prob.gradient()
    M = DiagonalPreconditioner(prob.fisher_diag(), DAMPING)

    def solve():
        return pcg_device(prob.matvec, -g, minv=M.minv, damping=DAMPING, max_iter=K_CG, tol=0.0, martens_conv_crit=False,
                          store_x_at_iters=None, poll=K_CG)
    t_solve = _timed(solve, flush, reps=10)
    v, out = torch.randn_like(theta), torch.empty_like(theta)
    t_mv = _timed(lambda: prob.matvec(v, out), flush)
    f_mv = flops_per_product(MLP_WIDTHS, MLP_BATCH)
    roof = contraction_roofline(lib, dev, pk, engine, MLP_BATCH, 512, 784, "layer-1 R-op forward", flush)
    api = api_steps(build_mlp(0).to(dev), loss_fn, [(xh.pin_memory(), th.pin_memory())], (xh, th), 5, None)
    return dict(workload="mlp_784-512-512-10_relu_ce_batch4096_ggn_fisher_pcg", params=theta.numel(),
                value=K_CG / (t_solve * 1e-3), unit=UNIT, ms_per_step=t_solve,
                matvec=dict(us_per_product=1e3 * t_mv, tflops_algorithmic=f_mv / (t_mv * 1e-3) / 1e12),
                roofline=roof, e2e_api=api)


def build_allcnnc(classes=100, seed=0):
    """All-CNN-C on a 3x32x32 input (the DeepOBS cifar100_allcnnc architecture of BASELINE.json configs[4], symmetric
    padding, eval mode: no dropout)."""
    nn = torch.nn
    torch.manual_seed(seed)

    def block(cin, cout, k, stride=1, pad=0):
        return [nn.Conv2d(cin, cout, k, stride=stride, padding=pad), nn.ReLU()]
    layers = (block(3, 96, 3, pad=1) + block(96, 96, 3, pad=1) + block(96, 96, 3, stride=2, pad=1) + block(96, 192, 3, pad=1)
              + block(192, 192, 3, pad=1) + block(192, 192, 3, stride=2, pad=1) + block(192, 192, 3) + block(192, 192, 1)
              + block(192, classes, 1) + [nn.AvgPool2d(6), nn.Flatten()])
    return nn.Sequential(*layers)


def conv_block(dev, engine, batch=1024):
    """BASELINE.json configs[4] (first conv slice): All-CNN-C, CrossEntropy, one GPU's share (1024) of the 8192 batch:
    time of one GGN-vector and one Hessian-vector product, each alone on the device.  F_Gv = 2N(4S - 2 m1),
    F_Hv = 2N(6S - 4 m1) with S = 271 420 416 MAC/sample, m1 = 2 654 208."""
    from pytorchhessianfree_b200.lowering import lower_module
    from pytorchhessianfree_b200.native import NativeNet
    from pytorchhessianfree_b200.problem import NativeProblem

    S, m1 = 271420416, 2654208
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    model = build_allcnnc().to(dev)
    loss_fn = torch.nn.CrossEntropyLoss()
    params = list(model.parameters())
    g = torch.Generator().manual_seed(5)
    x = torch.rand(batch, 3, 32, 32, generator=g).to(dev)
    t = torch.randint(0, 100, (batch,), generator=g).to(dev)
    prog = lower_module(model, loss_fn, params, input_shape=(3, 32, 32))
    theta = torch.cat([p.detach().reshape(-1) for p in params]).contiguous()
    net = NativeNet(prog.layers, prog.loss, prog.reduction, prog.n_params, engine=engine)
    out = dict(workload=f"allcnnc_3x32x32_cifar100_ce_batch{batch}_eval", params=theta.numel(), batch=batch)
    v, res = torch.randn_like(theta), torch.empty_like(theta)
    for curv, flops in (("ggn", 2.0 * batch * (4 * S - 2 * m1)), ("hessian", 2.0 * batch * (6 * S - 4 * m1))):
        prob = NativeProblem(net, theta, curv, [(x, t)])
        t_lin = _timed(prob.linearize, flush, reps=3)
        if curv == "hessian":
            prob.gradient()
        t_mv = _timed(lambda: prob.matvec(v, res), flush, reps=5)
        out[curv] = dict(ms_per_product=t_mv, products_per_s=1e3 / t_mv, tflops_algorithmic=flops / (t_mv * 1e-3) / 1e12,
                         ms_linearise=t_lin, workspace_gib=sum(l.workspace.numel() for l in prob.mvp_lins) / 2 ** 30)
        del prob
        torch.cuda.empty_cache()
    out["note"] = ("CPU reference (SURVEY.md section 6, 8 host cores, batch 64): 537 ms per _Gv, 1 282 ms per _Hv, i.e. 8.6 / 20.5 s at "
                   "this batch if it scaled linearly")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--engine", default=os.environ.get("HF_ENGINE", "tc"), choices=["simt", "tc"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
