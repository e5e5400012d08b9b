"""Drop-in ``HessianFree`` optimizer whose Newton-step solve runs on hand-written sm_100a kernels.

Public surface = the reference's (``hessianfree/optimizer.py``): constructor arguments and their
validation (``:23-115``), ``step`` (``:126-363``), ``acc_step`` (``:519-606``), ``test_reduction``
(``:817-926``), ``get_preconditioner`` (``:928-952``), the string-keyed ``state`` lists and the
``param_groups[0]`` keys.  What differs underneath:

* the curvature products (``_Gv``/``_Hv`` -> BackPACK -> autograd in the reference) are lowered once per
  step to a layer program and executed by ``hf_ggn_matvec`` / ``hf_hessian_matvec`` on activations that
  stay resident in HBM;
* the CG loop is :func:`pytorchhessianfree_b200.cg.pcg_device`: one fused vector kernel per iteration,
  termination decided on the device, no host sync per iteration;
* the loss evaluations of the LM ratio, cg-backtracking and the line search run as batched device passes;
* ``acc_step`` chunks may be sharded over the ranks of a ``torch.distributed`` process group
  (constructor argument ``process_group``): one all-reduce of the flat vector per CG iteration.

Two deliberate deviations: ``get_preconditioner`` *returns* the preconditioner
(the reference forgets the ``return``, ``optimizer.py:943``), and ``martens_conv_crit`` is accepted as a
keyword (default True = the reference's hard-wired behaviour, ``optimizer.py:271``).

The trainable parameters live in one flat FP32 buffer (``param.data`` are views into it) -- the layout
the reference ends up with after its first step (``utils.py:32``), made explicit.
"""
from contextlib import nullcontext
from warnings import warn

import torch
from torch.nn.utils.convert_parameters import parameters_to_vector

from . import _lib
from .cg import DiagonalPreconditioner, cg, pcg_device
from .cg_backtracking import cg_efficient_backtracking
from .linesearch import simple_linesearch
from .lowering import lower_graph, lower_module
from .native import NativeNet
from .problem import NativeProblem
from .utils import vector_to_trainparams


def _sample_shape(datalist):
    """(c, h, w) of one sample when the chunks hold image batches [n, c, h, w] (a convolution comes first), else None."""
    x = datalist[0][0]
    return tuple(x.shape[1:]) if x.dim() == 4 else None


class HessianFree(torch.optim.Optimizer):
    """Hessian-free optimizer (Martens 2010; Martens & Sutskever 2012)."""

    def __init__(
        self,
        params,
        curvature_opt="ggn",
        damping=1.0,
        adapt_damping=True,
        cg_max_iter=250,
        cg_decay_x0=0.95,
        use_cg_backtracking=True,
        lr=1.0,
        use_linesearch=True,
        verbose=False,
        *,
        martens_conv_crit=True,
        engine="auto",
        process_group=None,
    ):
        """Arguments up to ``verbose`` are the reference's (``optimizer.py:23-77``).

        Keyword-only extensions: ``martens_conv_crit`` (use Martens' relative-progress stopping rule in CG),
        ``engine`` ("auto" | "tc" | "simt": split-precision (TF32 + BF16 corrections) tensor-core tiles where the layer shapes allow, or FP32
        SIMT tiles everywhere), ``process_group`` (data-parallel group over which ``acc_step`` chunks are
        sharded; None = single GPU).
        """
        if curvature_opt not in ["hessian", "ggn"]:
            raise ValueError(f"Invalid curvature_opt = {curvature_opt}")
        if damping < 0.0:
            raise ValueError(f"Invalid damping = {damping}")
        self.adapt_damping = adapt_damping
        if damping == 0.0 and adapt_damping:
            self.adapt_damping = False
            warn("The damping is set to `0.0` and won't get adapted.")
        if cg_max_iter is not None and cg_max_iter < 1:
            raise ValueError(f"Invalid cg_max_iter: {cg_max_iter}")
        if lr < 0.0:
            raise ValueError(f"Invalid learning rate lr = {lr}")
        if engine not in ("auto", "tc", "simt"):
            raise ValueError(f"Invalid engine = {engine}")
        self.cg_decay_x0 = cg_decay_x0
        self.use_cg_backtracking = use_cg_backtracking
        self.use_linesearch = use_linesearch
        self.martens_conv_crit = martens_conv_crit
        self.engine = engine
        self.process_group = process_group

        super().__init__(params, dict(curvature_opt=curvature_opt, damping=damping, cg_max_iter=cg_max_iter, lr=lr))
        if len(self.param_groups) != 1:
            raise ValueError("`HessianFree` does not support per-parameter options.")
        self.verbose = verbose
        self._params = self._group["params"]
        self._params_list = [p for p in self._params if p.requires_grad]  # the subspace everything lives in
        self.device = self._params_list[0].device
        self._theta = None
        self._nets = {}

    @property
    def _group(self):
        """The one parameter group.  Looked up on every use: ``load_state_dict`` replaces the group dicts, and an
        alias taken in ``__init__`` (as in the reference, ``optimizer.py:118``) would keep reading the old damping."""
        return self.param_groups[0]

    # ------------------------------------------------------------------------------------------------
    # flat parameter buffer
    # ------------------------------------------------------------------------------------------------
    def _flat_params(self):
        """The flat FP32 vector the trainable parameters are views of (re-flattened if the user re-pointed
        a parameter since the last step)."""
        _lib.require_cuda(self._params_list[0], "the model parameters")
        th, off, ok = self._theta, 0, self._theta is not None
        if ok:
            for p in self._params_list:
                if p.data.data_ptr() != th.data_ptr() + 4 * off or p.dtype != torch.float32:
                    ok = False
                    break
                off += p.numel()
        if not ok:
            if any(p.dtype != torch.float32 for p in self._params_list):
                raise NotImplementedError("the native curvature kernels take float32 parameters")
            th = parameters_to_vector(self._params_list).detach().clone()
            vector_to_trainparams(th, self._params)
            self._theta = th
        return self._theta

    def _net_for(self, prog):
        engine = "tc" if self.engine in ("auto", "tc") else "simt"
        net = NativeNet(prog.layers, prog.loss, prog.reduction, prog.n_params, engine=engine)
        key = net.signature()
        return self._nets.setdefault(key, net)

    def _log_state(self):
        st = self.state
        st.setdefault("x0", None)
        for k in ("init_losses", "final_losses", "dampings", "cg_reasons", "num_cg_iters", "best_cg_iters",
                  "learning_rates"):
            st.setdefault(k, [])
        return st

    # ------------------------------------------------------------------------------------------------
    # step
    # ------------------------------------------------------------------------------------------------
    def step(self, forward, grad=None, mvp=None, M_func=None, test_deterministic=False):
        """One parameter update.  ``forward() -> (loss, outputs)``; ``grad``, ``mvp``, ``M_func`` are the
        reference's optional plug-ins (``optimizer.py:134-180``).  Without ``mvp`` the autograd graph of
        ``loss`` is lowered to the device layer program; an un-lowerable graph raises NotImplementedError."""
        if self.verbose:
            print("\nInformation on parameters...")
            print("  Total number of parameters: ", sum(p.numel() for p in self._params))
            print("  Number of trainable parameters: ", sum(p.numel() for p in self._params_list))
            print("  Device = ", self.device)
        if test_deterministic:
            self._test_forward_determinisitc(forward)
        with torch.no_grad() if (grad is not None and mvp is not None) else nullcontext():
            loss, outputs = forward()
        init_loss = loss.item()

        problem = None
        if mvp is None:
            ggn = self._group["curvature_opt"] == "ggn"
            if ggn and outputs is None:
                raise ValueError("curvature_opt='ggn' needs the `outputs` returned by `forward`")
            prog = lower_graph(loss, outputs if ggn else None, self._params_list)
            theta = self._flat_params()
            problem = NativeProblem(self._net_for(prog), theta, self._group["curvature_opt"],
                                    [(prog.inputs, prog.targets)], group=None)
            native_loss = float(problem.linearize().item())
            if abs(native_loss - init_loss) > 1e-4 * max(1.0, abs(init_loss)):
                raise RuntimeError(f"the lowered layer program evaluates to {native_loss:.6e} but `forward` returned "
                                   f"{init_loss:.6e}: the model contains something the lowering did not see")
            mvp_plain = problem.mvp
        else:
            mvp_plain = mvp
        if grad is None:
            if problem is not None:
                grad = problem.gradient()
            else:
                g = torch.autograd.grad(loss, self._params_list, create_graph=True, retain_graph=True)
                grad = parameters_to_vector(g).detach()
        elif problem is not None and self._group["curvature_opt"] == "hessian":
            problem.gradient()  # the Hessian product needs the back-propagated deltas
        if test_deterministic:
            self._test_mvp_deterministic(mvp_plain)
        return self._newton_step(init_loss, grad, problem, mvp_plain, M_func, forward)

    def _newton_step(self, init_loss, grad, problem, mvp, M_func, forward):
        """Solve (B + lambda I) x = -grad, pick the step, update (reference ``optimizer.py:253-363``)."""
        state = self._log_state()
        if self.verbose:
            print(f"\nInitial loss = {init_loss:.6f}")
        state["init_losses"].append(init_loss)
        damping = self._group["damping"]
        state["dampings"].append(damping)
        grid = None if self.use_cg_backtracking else [0]
        b = -grad

        fused_M = M_func is None or isinstance(M_func, DiagonalPreconditioner)
        if problem is not None and fused_M:
            x_iters, m_iters, cg_reason = pcg_device(
                problem.matvec, b, x0=state["x0"], minv=None if M_func is None else M_func.minv, damping=damping,
                max_iter=self._group["cg_max_iter"], martens_conv_crit=self.martens_conv_crit,
                store_x_at_iters=grid, verbose=self.verbose, out_buffer=problem.out_buffer())
        else:
            x_iters, m_iters, cg_reason = cg(
                A=lambda x: mvp(x) + damping * x, b=b, x0=state["x0"], M=M_func, max_iter=self._group["cg_max_iter"],
                martens_conv_crit=self.martens_conv_crit, store_x_at_iters=grid, verbose=self.verbose)
        state["cg_reasons"].append(cg_reason)
        state["num_cg_iters"].append(len(x_iters) - 1)
        step_vec = x_iters[-1]
        self._set_x0(self.cg_decay_x0 * x_iters[-1])  # the un-backtracked solution, Martens 2010 sec. 4.6

        f_at_zero = None
        if problem is not None:
            params_vec = problem.theta
            tfunc = problem.target_function()
            # one device pass and one host sync for everything the rest of the step usually asks: f(0) for the line
            # search, f(x_0) and f(x_last) for the damping ratio, the first lookahead window of the backtracking walk
            zero = torch.zeros_like(step_vec)
            ask = [zero]
            if self.adapt_damping:
                ask += [x_iters[0], x_iters[-1]]
            if self.use_cg_backtracking:
                ask += [s for s in reversed(x_iters) if s is not None][:3]
            if not self.use_linesearch:
                ask = ask[1:]
            primed = tfunc.prime(ask) if ask else []
            f_at_zero = primed[0] if self.use_linesearch else None
        else:
            params_vec = parameters_to_vector(self._params_list).detach()

            @torch.no_grad()
            def tfunc(step):
                vector_to_trainparams(params_vec + step, self._params)
                return forward()[0].item()

        assert x_iters[0] is not None and x_iters[-1] is not None
        if self.adapt_damping:
            if m_iters is None:
                raise ValueError("adapt_damping needs the quadratic-model values: keep martens_conv_crit=True")
            many = getattr(tfunc, "many", None)
            f_0, f_step = many([x_iters[0], x_iters[-1]]) if many else (tfunc(x_iters[0]), tfunc(x_iters[-1]))
            self._adapt_damping(f_0=f_0, f_step=f_step, m_0=m_iters[0].item(), m_step=m_iters[-1].item())

        if self.use_cg_backtracking:
            best_cg_iter, _ = cg_efficient_backtracking(f=tfunc, steps_list=x_iters, verbose=self.verbose,
                                                        lookahead=3 if problem is not None else 1)
            state["best_cg_iters"].append(best_cg_iter)
            step_vec = x_iters[best_cg_iter]

        lr = self._group["lr"]
        if not self.use_linesearch:
            if self.verbose:
                print(f"\nConstant lr = {lr:.6f}")
            final_loss = None
        else:
            lr, final_loss = simple_linesearch(f=tfunc, f_grad_0=grad, step=step_vec, init_alpha=lr, verbose=self.verbose,
                                               f_0=f_at_zero)
        state["learning_rates"].append(lr)

        if self.verbose:
            print(f"\nParameter update with lr = {lr:.6f}")
        if problem is not None:
            params_vec.add_(step_vec, alpha=lr)  # parameters are views of this buffer
        else:
            vector_to_trainparams(params_vec + lr * step_vec, self._params)
            self._theta = None
        if self.verbose:
            if final_loss is None:
                final_loss = tfunc(torch.zeros_like(step_vec)) if problem is not None else forward()[0].item()
            state["final_losses"].append(final_loss)
            print(f"Initial loss = {init_loss:.6f} --> final loss = {final_loss:.6f}")
        return final_loss

    # ------------------------------------------------------------------------------------------------
    # determinism self-tests (reference optimizer.py:365-448)
    # ------------------------------------------------------------------------------------------------
    def _test_forward_determinisitc(self, forward):
        """Two calls of ``forward`` must agree; warns otherwise (name kept from the reference)."""
        l1, o1 = forward()
        l2, o2 = forward()
        same = torch.allclose(l1, l2)
        if o1 is not None and o2 is not None:
            same = same and torch.allclose(o1, o2)
        if self.verbose:
            print("\nTest deterministic behavior of `forward`: " + ("passed" if same else "failed"))
        if not same:
            warn("Non-determinisitc behaviour detected. Consider setting your model to evaluation mode, "
                 "i.e. `model.eval()`.")

    def _test_mvp_deterministic(self, mvp):
        """Two products with the same random vector must agree; warns otherwise."""
        x = torch.randn_like(parameters_to_vector(self._params_list)).to(self.device)
        same = torch.allclose(mvp(x), mvp(x))
        if self.verbose:
            print("\nTest deterministic behavior of `mvp`: " + ("passed" if same else "failed"))
        if not same:
            warn("Non-determinisitc behaviour detected. Consider setting your model to evaluation mode, "
                 "i.e. `model.eval()`.")

    # ------------------------------------------------------------------------------------------------
    # curvature products with the reference's static signatures (optimizer.py:450-462)
    # ------------------------------------------------------------------------------------------------
    @staticmethod
    def _native_product(loss, outputs, params_list, vec, curvature_opt):
        prog = lower_graph(loss, outputs, params_list)
        theta = parameters_to_vector(params_list).detach().to(torch.float32)
        net = NativeNet(prog.layers, prog.loss, prog.reduction, prog.n_params)
        problem = NativeProblem(net, theta, curvature_opt, [(prog.inputs, prog.targets)])
        problem.linearize()
        if curvature_opt == "hessian":
            problem.gradient()
        return problem.mvp(vec)

    @staticmethod
    def _Hv(loss, params_list, vec):
        """Hessian-vector product of ``loss`` w.r.t. ``params_list`` (flat in, flat out)."""
        return HessianFree._native_product(loss, None, params_list, vec, "hessian")

    @staticmethod
    def _Gv(loss, outputs, params_list, vec):
        """GGN-vector product  J^T H_loss J v  (flat in, flat out)."""
        return HessianFree._native_product(loss, outputs, params_list, vec, "ggn")

    def _adapt_damping(self, f_0, f_step, m_0, m_step):
        """Levenberg-Marquardt update of the damping from the reduction ratio (reference ``:464-506``):
        rho < 1/4 -> x3/2, rho > 3/4 -> x2/3; warns when rho < 0."""
        rho = (f_step - f_0) / (m_step - m_0)
        if self.verbose:
            print("\nLM-heurisitc: Adapt damping...")
            print(f"  f_0 = {f_0:.6f}, f_step = {f_step:.6f}, m_0 = {m_0:.6f}, m_step = {m_step:.6f}")
            print(f"  Reduction ratio rho = {rho:.6f}")
        if rho < 0.25:
            self._group["damping"] *= 3 / 2
        elif rho > 0.75:
            self._group["damping"] *= 2 / 3
        if self.verbose:
            print(f"  Damping is set to {self._group['damping']:.6f}")
        if rho < 0:
            warn("The reduction ratio `rho` is negative. This might result in a bad cg-initialization in the "
                 "next step.")

    def _set_x0(self, new_x0):
        self.state["x0"] = new_x0

    # ------------------------------------------------------------------------------------------------
    # acc_step
    # ------------------------------------------------------------------------------------------------
    def acc_step(self, model, loss_func, loss_datalist, grad_datalist=None, mvp_datalist=None, M_func=None,
                 reduction="mean", test_deterministic=False):
        """Step with loss, gradient and curvature each accumulated over a list of ``(inputs, targets)``
        chunks (reference ``optimizer.py:519-606``).  With a ``process_group`` the lists are this rank's
        shard of the global lists; sums run over all ranks."""
        if reduction not in ["mean", "sum"]:
            raise ValueError(f"Invalid reduction {reduction}")
        prog = lower_module(model, loss_func, self._params_list, input_shape=_sample_shape(mvp_datalist or loss_datalist))
        if prog.reduction != reduction:
            raise ValueError(f"the loss function reduces by {prog.reduction!r} but reduction={reduction!r} was given "
                             "(check with `test_reduction`)")
        theta = self._flat_params()
        dev = lambda dl: [(x.to(self.device), t.to(self.device)) for x, t in dl]  # noqa: E731
        grad_datalist = loss_datalist if grad_datalist is None else grad_datalist  # reference :575-579
        mvp_datalist = loss_datalist if mvp_datalist is None else mvp_datalist
        net, mvp_data = self._net_for(prog), dev(mvp_datalist)
        cached, self._prelinearized = getattr(self, "_prelinearized", None), None
        same_lists = grad_datalist is mvp_datalist and loss_datalist is mvp_datalist
        if (cached is not None and same_lists and self._group["curvature_opt"] == "ggn"
                and cached[0] == self._problem_key(net, theta, mvp_data) and torch.equal(cached[3], theta)):
            _, problem, mvp_loss, _ = cached  # linearised by get_preconditioner on the same data at the same parameters
        else:
            problem = NativeProblem(
                net, theta, self._group["curvature_opt"], mvp_data=mvp_data,
                grad_data=None if grad_datalist is mvp_datalist else dev(grad_datalist),
                loss_data=None if loss_datalist is mvp_datalist else dev(loss_datalist),
                group=self.process_group)
            mvp_loss = problem.linearize()
        grad = problem.gradient()
        if loss_datalist is mvp_datalist:
            init_loss = float(mvp_loss.to(torch.float32).item())
        else:
            init_loss = problem.losses_at([torch.zeros_like(theta)])[0]
        if test_deterministic:
            self._test_mvp_deterministic(problem.mvp)
        return self._newton_step(init_loss, grad, problem, problem.mvp, M_func, None)

    # ------------------------------------------------------------------------------------------------
    # misc
    # ------------------------------------------------------------------------------------------------
    def test_reduction(self, model, loss_func, datalist, reduction):
        """Check that ``reduction`` matches the loss function: loss, gradient and curvature product
        accumulated over ``datalist`` with ``reduction`` must equal the ones of the concatenated batch
        (reference ``optimizer.py:817-926``; rtol 1e-2, atol 1e-4).  Raises RuntimeError otherwise."""
        if reduction not in ["mean", "sum"]:
            raise ValueError(f"Invalid reduction {reduction}")
        assert len(datalist) > 1, "This test is only meaningful for a data list with at least two entries."
        prog = lower_module(model, loss_func, self._params_list, input_shape=_sample_shape(datalist))
        theta = self._flat_params()
        net = self._net_for(prog)
        curv = self._group["curvature_opt"]
        x = torch.randn_like(theta)

        def quantities(data):
            prob = NativeProblem(net, theta, curv, [(a.to(self.device), t.to(self.device)) for a, t in data])
            loss = prob.linearize().to(torch.float32)
            g = prob.gradient()
            return loss, g, prob.mvp(x), prob.n_mvp

        # chunk-wise with the loss function's own reduction, then weighted as `reduction` says (:678-684)
        acc, total = None, 0
        for chunk in datalist:
            l, g, m, n = quantities([chunk])
            w = float(n) if reduction == "mean" else 1.0
            acc = [w * l, w * g, w * m] if acc is None else [a + w * q for a, q in zip(acc, (l, g, m))]
            total += n
        if reduction == "mean":
            acc = [a / total for a in acc]
        ref = quantities([(torch.cat([a for a, _ in datalist]), torch.cat([t for _, t in datalist]))])[:3]
        ok = True
        for name, r, a in zip(("loss values", "gradients", "mvps"), ref, acc):
            good = torch.allclose(a, r, rtol=1e-2, atol=1e-4)
            if self.verbose:
                print(f"  Test {name}: " + ("passed" if good else "failed"))
            ok = ok and good
        if not ok:
            raise RuntimeError(f"Inconsistent results for reduction {reduction}. This could also be the result of "
                               "non-deterministic behavior or simply due to using the GPU.")
        if self.verbose:
            print("  All tests passed")

    def get_preconditioner(self, model, loss_func, inputs, targets, reduction, exponent=None, use_backpack=True):
        """Empirical-Fisher diagonal preconditioner with the optimizer's current damping
        (reference ``optimizer.py:928-952``).  Unlike the reference, the preconditioner is returned."""
        if reduction not in ["sum", "mean"]:
            raise ValueError(f"reduction {reduction} is not supported.")
        prog = lower_module(model, loss_func, self._params_list, input_shape=_sample_shape([(inputs, targets)]))
        if prog.reduction != reduction:
            raise ValueError(f"the loss function reduces by {prog.reduction!r} but reduction={reduction!r} was given")
        theta = self._flat_params()
        data = [(inputs.to(self.device), targets.to(self.device))]
        net = self._net_for(prog)
        problem = NativeProblem(net, theta, "ggn", data, group=self.process_group)
        loss = problem.linearize()
        diag = problem.fisher_diag()
        # The usual call order is get_preconditioner(x, t) followed by acc_step([(x, t)]) at the same parameters: keep
        # the linearisation so that acc_step does not repeat the forward pass (dropped as soon as anything differs).
        # The key holds addresses and version counters; the parameter VALUES are compared too (a snapshot of theta,
        # one device-side equality test in acc_step): `param.data` are views re-pointed into theta, so an in-place
        # update of a parameter or `load_state_dict` changes theta without touching theta's own version counter.
        self._prelinearized = (self._problem_key(net, theta, data), problem, loss, theta.clone())
        return DiagonalPreconditioner(diag, self._group["damping"], 0.75 if exponent is None else exponent)

    def _problem_key(self, net, theta, data):
        """Identity of a linearisation: net, parameter buffer, the version counters of the buffer and of every
        parameter viewing it, data buffers and their versions (data edited through ``.data`` escapes the counters:
        pass a new tensor, or call ``acc_step`` without a preceding ``get_preconditioner`` on that data)."""
        return (net.signature(), theta.data_ptr(), theta._version, tuple(p._version for p in self._params_list),
                tuple((x.data_ptr(), x._version, tuple(x.shape), t.data_ptr(), t._version, tuple(t.shape)) for x, t in data))
