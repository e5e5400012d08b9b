"""CG-backtracking: pick the snapshot of the solve with the lowest loss (Martens 2010, sec. 4.6;
reference ``hessianfree/cg_backtracking.py``).  ``f`` may evaluate a whole list of candidates in one
device pass (``f.many``, provided by the native optimizer path); the selection rule is unchanged."""
import torch


def _evaluate(f, steps, idxs):
    many = getattr(f, "many", None)
    if many is not None and len(idxs) > 1:
        return many([steps[i] for i in idxs])
    return [f(steps[i]) for i in idxs]


def _report(steps, vals, best):
    for i, v in enumerate(vals):
        if steps[i] is None:
            continue
        if v is None:
            print(f"  cg-iteration {i}, loss not evaluated")
        else:
            print(("* " if i == best else "  ") + f"cg-iteration {i}, loss = {v:.6f}")


def cg_backtracking(f, steps_list, verbose=False):
    """Evaluate every candidate; return ``(index of the minimum, minimum)`` (reference ``:6-50``)."""
    if verbose:
        print("\nBacktracking cg-iterations...")
    live = [i for i, s in enumerate(steps_list) if s is not None]
    got = dict(zip(live, _evaluate(f, steps_list, live)))
    vals = [got.get(i, float("inf")) for i in range(len(steps_list))]
    best = int(torch.argmin(torch.tensor(vals, dtype=torch.float32)))
    if verbose:
        _report(steps_list, [got.get(i) for i in range(len(steps_list))], best)
    return best, vals[best]


def cg_efficient_backtracking(f, steps_list, verbose=False, lookahead=1):
    """Walk back from the last candidate while the loss keeps improving (reference ``:53-112``).

    ``lookahead`` candidates are evaluated per device pass; the extra evaluations never change the
    answer because the walk still stops at the first candidate that fails to improve.
    """
    if verbose:
        print("\nBacktracking cg-iterations...")
    live = [i for i in range(len(steps_list) - 1, -1, -1) if steps_list[i] is not None]
    seen = [None] * len(steps_list)
    best, best_val, stop = None, float("inf"), False
    for lo in range(0, len(live), max(1, lookahead)):
        chunk = live[lo: lo + max(1, lookahead)]
        for i, v in zip(chunk, _evaluate(f, steps_list, chunk)):
            seen[i] = v
            if v < best_val:
                best, best_val = i, v
            else:
                stop = True
                break
        if stop:
            break
    if best is None:
        raise RuntimeError("cg-backtracking found no finite loss among the candidates")
    if verbose:
        _report(steps_list, seen, best)
    return best, best_val
