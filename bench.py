"""Benchmark of the hot path: fixed-K preconditioned-CG solves with GGN-vector products on the
BASELINE.json configs[1] workload (MLP 784-512-512-10 ReLU, CrossEntropy, batch 4096 per GPU,
Fisher-diagonal PCG).

    python bench.py --gpus 1 --steps 20 --warmup 3            # this repo's sm_100a path
    python bench.py --impl reference --steps 5 --warmup 1     # the CPU oracle port of the reference path
    torchrun --nproc-per-node N ... bench.py --gpus N ...     # batch sharded over N ranks (weak scaling)

One "step" = one PCG solve of (G + lambda I) x = -g with exactly K_cg = 50 iterations (tol = 0, Martens'
criterion off, so only the iteration cap stops it -- BASELINE.md section 3; all termination quantities are
still computed every iteration).  Each iteration = one GGN-vector product over the rank's 4096-sample
shard (+ one all-reduce of the P-vector when N > 1) + one fused CG vector update.  The unit counted is
that per-shard product, so `value` = steps * K_cg * N / seconds.

Prints ONE JSON line (rank 0).  Timing: CUDA events on the launching stream around every step, max over
ranks, L2 flushed (256 MiB write) between steps outside the timed region.
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

WIDTHS = [784, 512, 512, 10]
BATCH = 4096
K_CG = 50
DAMPING = 1e-3  # small enough that 50 iterations stay clear of the float32 rounding floor (with lambda = 1 the
#                 solve converges in ~13 iterations and a fixed-50 run would divide 0/0)
METRIC = "GGN-vector products/sec (CG iters/sec)"
UNIT = "products/s"


def build_mlp(seed=0):
    torch.manual_seed(seed)
    mods = []
    for i in range(len(WIDTHS) - 1):
        mods.append(torch.nn.Linear(WIDTHS[i], WIDTHS[i + 1]))
        if i < len(WIDTHS) - 2:
            mods.append(torch.nn.ReLU())
    return torch.nn.Sequential(*mods)


def synth(seed, n=BATCH):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n, WIDTHS[0], generator=g), torch.randint(0, WIDTHS[-1], (n,), generator=g)


def flops_per_product(n=BATCH):
    """F_Gv = 2N(4S - 2 m1): contractions only."""
    m = [WIDTHS[i] * WIDTHS[i + 1] for i in range(len(WIDTHS) - 1)]
    return 2.0 * n * (4 * sum(m) - 2 * m[0])


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained"), src="measured")
    return dict(hbm=6650.0, tf=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        busy = [v for v in sm if v > 0.5 * mx] or sm
        return dict(sm_mhz=busy[len(busy) // 2] if busy else None, sm_max_mhz=mx or None, reasons=sorted(reasons),
                    samples=len(sm))


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path (reference cg.py + _Gv through the BackPACK recipe)
# ------------------------------------------------------------------------------------------------------
def cpu_solve_rate(solves, k_cg, threads=None):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import hf_oracle as O

    if threads:
        torch.set_num_threads(threads)
    model = build_mlp(0)
    loss_fn = torch.nn.CrossEntropyLoss()
    x, t = synth(1)
    params = list(model.parameters())
    out = model(x)
    loss = loss_fn(out, t)
    grad = O.flatten(torch.autograd.grad(loss, params, create_graph=True)).detach()
    M = O.diag_precond(torch.rand_like(grad) * 1e-3, DAMPING)  # same kind of operator; values do not affect cost
    A = lambda v: O.Gv(loss, out, params, v) + DAMPING * v  # noqa: E731
    O.pcg(A, -grad, M=M, max_iter=2, tol=0.0)  # warm-up
    t0 = time.perf_counter()
    for _ in range(solves):
        O.pcg(A, -grad, M=M, max_iter=k_cg, tol=0.0)
    dt = time.perf_counter() - t0
    # the reference's cg spends K+1 products for K iterations (cg.py:188)
    return solves * k_cg / dt, dt, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    k_cg = 5  # bounded sample: 5 CG iterations per step instead of 50, identical per-iteration work
    cpu_solve_rate(max(1, args.warmup), k_cg)
    rate, dt, threads = cpu_solve_rate(args.steps, k_cg)
    line = dict(metric=METRIC, value=rate, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * dt / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", impl="reference",
                config=dict(workload="mlp_784-512-512-10_relu_ce_batch4096_ggn_fisher_pcg", cg_iters_per_step=k_cg,
                            damping=DAMPING, note="oracle port of reference cg.py + _Gv (BackPACK autograd recipe) on host cores"),
                cpu_baseline=dict(value=rate, unit=UNIT, cores=threads, kind="port",
                                  sample=f"{args.steps} solves x {k_cg} CG iterations, batch {BATCH}"),
                e2e=dict(value=rate, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def run_native(args):
    import torch.distributed as dist

    from pytorchhessianfree_b200 import DiagonalPreconditioner, _lib, pcg_device
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

    model = build_mlp(0).to(dev)
    loss_fn = torch.nn.CrossEntropyLoss()
    params = list(model.parameters())
    prog = lower_module(model, loss_fn, params)
    theta = torch.cat([p.detach().reshape(-1) for p in params]).contiguous()
    P = theta.numel()
    net = NativeNet(prog.layers, prog.loss, prog.reduction, prog.n_params, engine=engine)
    x_host, t_host = synth(1 + rank)
    x_host, t_host = x_host.pin_memory(), t_host.pin_memory()
    x, t = x_host.to(dev), t_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def setup(xd, td):
        prob = NativeProblem(net, theta, "ggn", [(xd, td)], group=group)
        prob.linearize()
        g = prob.gradient()
        M = DiagonalPreconditioner(prob.fisher_diag(), DAMPING)
        return prob, g, M

    def solve(prob, g, M):
        return pcg_device(prob.matvec, -g, minv=M.minv, damping=DAMPING, max_iter=K_CG, tol=0.0,
                          martens_conv_crit=False, store_x_at_iters=None, poll=K_CG)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm: inputs in HBM, linearisation done, time the solves -------------------
    prob, g, M = setup(x, t)
    for _ in range(args.warmup):
        solve(prob, g, M)
    barrier()
    n0 = lib.hf_debug_launch_count()
    times = []
    with ClockSampler(local) as clk:
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
        # nvidia-smi reports every ~100 ms and a 10-step timed region lasts ~0.1 s: keep the identical load running
        # (untimed, same collectives on every rank) until the sampler has seen it
        for _ in range(12):
            for _ in range(4):
                solve(prob, g, M)
            torch.cuda.synchronize()
    assert why == "Number of iterations" and len(xs) == K_CG + 1, f"fixed-K solve stopped early: {why}, {len(xs) - 1} iterations"
    total_ms = torch.tensor([sum(times)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    value = args.steps * K_CG * world / (total_ms * 1e-3)

    # ---- end-to-end arm: host buffers in, result out, everything inside the timed region -----------
    def e2e_step():
        xd, td = x_host.to(dev, non_blocking=True), t_host.to(dev, non_blocking=True)
        pr, gg, MM = setup(xd, td)
        xs_, _, _ = solve(pr, gg, MM)
        return xs_[-1].cpu()  # the Newton step (P floats) read back

    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    barrier()
    e2e_steps = max(3, args.steps // 2)
    e2e_ms = 0.0
    for _ in range(e2e_steps):
        flush.fill_(1)
        barrier()
        t0 = time.perf_counter()
        e2e_step()
        torch.cuda.synchronize()
        e2e_ms += 1e3 * (time.perf_counter() - t0)
    e2e_t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = e2e_steps * K_CG * world / (float(e2e_t.item()) * 1e-3)

    # ---- per-kernel rooflines, measured live (rank 0, kernels alone on the device) ------------------
    pk = peaks()
    roof, extra = None, {}
    if rank == 0:
        # rank-local problem: the per-kernel timings must not issue collectives the other ranks do not join
        local_prob = NativeProblem(net, theta, "ggn", [(x, t)], group=None)
        roof, extra = kernel_rooflines(lib, local_prob, theta, dev, pk, engine)

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            rate, dt, threads = cpu_solve_rate(2, 50)
            cpu = dict(value=rate, unit=UNIT, cores=threads, kind="port",
                       sample=f"2 solves x 50 CG iterations of the same workload ({dt:.1f} s)")
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=total_ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f32", data="synthetic",
                    config=dict(workload="mlp_784-512-512-10_relu_ce_batch4096_ggn_fisher_pcg", batch_per_gpu=BATCH,
                                params=P, cg_iters_per_step=K_CG, damping=DAMPING, engine=engine,
                                l2="flushed between steps (256 MiB write)",
                                parallelism=f"dp{world}: batch sharded, all-reduce of the P-vector per CG iteration"),
                    clocks=clk.summary(), gpu_launches=int(launches),
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=int(x_host.numel() * 4 + t_host.numel() * 8),
                             d2h_bytes_per_step=int(P * 4),
                             what="H2D inputs + forward + gradient + Fisher diagonal + 50-iteration PCG + D2H step"),
                    roofline=roof, cpu_baseline=cpu, **extra)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def kernel_rooflines(lib, prob, theta, dev, pk, engine):
    """Time (a) one GGN-vector product and its dominant contraction, (b) the fused CG vector update, each alone
    on the device with CUDA events, L2 flushed between launches."""
    from pytorchhessianfree_b200.cg import _Solver
    from pytorchhessianfree_b200._lib import PCG_FUSED

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn, reps=10):
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

    prob.linearize()
    v, out = torch.randn_like(theta), torch.empty_like(theta)
    t_mv = timed(lambda: prob.matvec(v, out))
    f_mv = flops_per_product()
    # dominant kernel: the tensor-tile contraction.  Timed on the shape that heads the launch list of a product, the
    # layer-1 R-op forward  Rz1[4096,512] = a0[4096,784] V1[512,784]^T  (128 CTAs = one wave), through the same C entry
    # point the products use.
    from pytorchhessianfree_b200._lib import Operand
    a0 = torch.randn(BATCH, 784, device=dev)
    v1 = torch.randn(512, 784, device=dev)
    c = torch.empty(BATCH, 512, device=dev)
    A = (Operand * 1)(Operand(a0.data_ptr(), 784, 1))
    B = (Operand * 1)(Operand(v1.data_ptr(), 784, 1))
    eng = 1 if engine == "tc" else 0
    stream = torch.cuda.current_stream().cuda_stream

    def dom():
        rc = lib.hf_contract(eng, BATCH, 512, 784, 1, A, B, c.data_ptr(), 512, None, 0, stream)
        assert rc == 0, lib.hf_last_error_string()
    t_dom = timed(dom)
    f_dom = 2.0 * BATCH * 512 * 784
    roof = dict(bound="tensor", achieved=f_dom / (t_dom * 1e-3) / 1e12, peak=pk["tf"], unit="TFLOP/s",
                frac=f_dom / (t_dom * 1e-3) / 1e12 / pk["tf"],
                traffic=17.9e6 if engine == "tc" else None,  # dram read+write per launch, ncu --set full (profiles/)
                peak_source=pk["src"],
                kernel=f"contraction 4096x512x784 (layer-1 R-op forward), engine={engine}",
                us_per_launch=1e3 * t_dom,
                note="split precision (TF32 main term + two BF16 correction terms = 4 bf16-MMA equivalents per product), "
                     "so the algorithmic ceiling is 1/4 of the bf16 peak")
    # fused CG vector update at the Martens-autoencoder size (P = 2,837,314: larger than cfg2 so that the pass is
    # bandwidth- rather than latency-bound) and at this workload's own P
    upd = {}
    for name, P in (("cfg2_P669706", theta.numel()), ("cfg3_P2837314", 2837314)):
        b = torch.randn(P, device=dev)
        s = _Solver(b, 10 ** 6)
        minv = torch.rand(P, device=dev) + 0.5
        Bp = torch.randn(P, device=dev)
        s.init(None, None, minv, DAMPING, 0.0, None, False, False)
        t_u = timed(lambda: s.iterate(PCG_FUSED, Bp=Bp, minv=minv, lam=DAMPING), reps=20)
        gbs = 36.0 * P / (t_u * 1e-3) / 1e9
        upd[name] = dict(bound="hbm", achieved=gbs, peak=pk["hbm"], unit="GB/s", frac=gbs / pk["hbm"],
                         us_per_launch=1e3 * t_u, bytes_per_launch=36 * P)
    extra = dict(roofline_cg_update=upd,
                 matvec=dict(us_per_product=1e3 * t_mv, tflops_algorithmic=f_mv / (t_mv * 1e-3) / 1e12,
                             flops_per_product=f_mv, frac_of_bf16_peak=f_mv / (t_mv * 1e-3) / 1e12 / pk["tf"]))
    return roof, extra


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
