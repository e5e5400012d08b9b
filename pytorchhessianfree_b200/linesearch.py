"""Armijo back-off line search (Martens & Sutskever 2012, sec. 8.8; reference ``hessianfree/linesearch.py``)."""
from warnings import warn

import torch


def simple_linesearch(f, f_grad_0, step, init_alpha=1.0, beta=0.8, c=1e-2, max_iter=20, verbose=False, f_0=None):
    """Shrink ``alpha`` by ``beta`` until ``f(alpha*step) <= f(0) + alpha*c*<grad, step>``.

    Returns ``(alpha, f(alpha*step))``; ``(0.0, f(0))`` with a warning when ``max_iter`` trials fail
    (reference ``linesearch.py:100-103``).  If ``f`` offers ``f.many`` the first two evaluations
    (``f(0)`` and ``f(init_alpha*step)``) share one device pass.  ``f_0`` (extension): the value of ``f(0)`` when the
    caller already has it.  With ``init_alpha == 1`` the candidate passed to ``f`` is ``step`` itself, not a copy,
    so that a memoising ``f`` recognises it.
    """
    if beta >= 1.0:
        raise ValueError(f"Invalid reduction factor beta = {beta}")
    if c < 0.0:
        raise ValueError(f"Invalid c = {c}")
    if verbose:
        print("\nStarting line search...")
    many = getattr(f, "many", None)
    first = step if init_alpha == 1.0 else init_alpha * step
    if f_0 is not None:
        f_0, f_try = float(f_0), float(f(first))
    elif many is not None:
        f_0, f_try = (float(v) for v in many([torch.zeros_like(step), first]))
    else:
        f_0 = float(f(torch.zeros_like(step)))
        f_try = float(f(first))
    if verbose:
        print(f"  f(0) = {f_0:.6f}")
        print(f"  f(init_alpha * step) = {f_try:.6f}")
    slope = c * torch.dot(f_grad_0, step).item()
    if slope >= 0:
        warn("`update_vec`-parameter in `simple_linesearch` is not a descent "
             f"direction. The directional derivative is {slope:.6f}.")
    alpha = init_alpha
    for _ in range(max_iter):
        if verbose:
            print(f"  Trying alpha = {alpha:.6f}, f(alpha * step) = {f_try:.6f}")
        if float(f_try) <= f_0 + alpha * slope:
            if verbose:
                print(f"Significant improvement for alpha = {alpha:.6f}")
            return alpha, f_try
        alpha *= beta
        f_try = f(alpha * step)
    warn("No suitable update could be found by the line search.")
    if verbose:
        print(f"No significant improvement. Using alpha = {0.0:.6f}")
    return 0.0, f_0
