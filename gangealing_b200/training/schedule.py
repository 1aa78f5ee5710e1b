"""psi annealing and the cyclic learning-rate schedule of the training loop -- host-side mirror of reference
utils/annealing.py and train.py:89-96,129-132, in closed form (plain floats: the Trainer copies them into the device
scalars its captured CUDA graph reads, so one graph serves the whole schedule).

  psi(i)            reference get_psi_annealing_fn('cosine'|'linear') evaluated while i <= anneal_psi, 0 afterwards
  lr(epoch)         reference DecayingCosineAnnealingWarmRestarts.step(epoch) with T_0 = 1 (train.py:206-207), where
                    epoch = max(0, (i - anneal_psi) / period) once psi is fully annealed (train.py:129-132)
"""
import math


def cosine_anneal(i, maxval, minval, num_steps):
    return minval + 0.5 * (maxval - minval) * (1 + math.cos(math.pi * i / num_steps))


def linear_anneal(i, maxval, minval, num_steps):
    return maxval - i * (maxval - minval) / num_steps


def fastslow_anneal(i, maxval, minval, num_steps, a=0.3):
    assert maxval == 1.0 and minval == 0.0
    na = num_steps * a
    return (na - a * i) / (na + i)


def get_psi_annealing_fn(anneal_fn):
    if anneal_fn == "linear":
        return linear_anneal
    if anneal_fn == "cosine":
        return cosine_anneal
    raise NotImplementedError(anneal_fn)


def psi_at(i, anneal_psi, anneal_fn="cosine"):
    """-> (psi, psi_is_fixed) at iteration i (reference train.py:91-96)."""
    if i <= anneal_psi:
        return float(get_psi_annealing_fn(anneal_fn)(i, 1.0, 0.0, anneal_psi)), False
    return 0.0, True


def decaying_cosine_lr(epoch, base_lr, t_mult=2, decay=0.9, eta_min=0.0, t_0=1):
    """Learning rate DecayingCosineAnnealingWarmRestarts assigns in `step(epoch)` (fractional epochs allowed)."""
    if epoch < 0:
        raise ValueError("Expected non-negative epoch, but got {}".format(epoch))
    if epoch >= t_0:
        if t_mult == 1:
            t_cur, t_i, n = epoch % t_0, t_0, int(epoch // t_0)
        else:
            n = int(math.log((epoch / t_0 * (t_mult - 1) + 1), t_mult))
            t_cur = epoch - t_0 * (t_mult ** n - 1) / (t_mult - 1)
            t_i = t_0 * t_mult ** n
    else:
        t_cur, t_i, n = epoch, t_0, 0
    return (decay ** n) * (eta_min + (base_lr - eta_min) * (1 + math.cos(math.pi * t_cur / t_i)) / 2)


def lr_cycle_iters(anneal_psi, period, iters, tm):
    """Iterations at which the learning rate reaches zero (reference annealing.py:44-51)."""
    zero_lr_iters = [anneal_psi - 1]
    num_cycles = int(math.log((iters - anneal_psi) / period, tm))
    for n in range(num_cycles):
        zero_lr_iters.append(int(zero_lr_iters[-1] + period * tm ** n))
    return zero_lr_iters


def schedule_at(i, stn_lr, ll_lr, anneal_psi=150000, period=37500, tm=2, decay=0.9, anneal_fn="cosine"):
    """-> dict(psi, stn_lr, ll_lr) for iteration i of the reference recipe (defaults: utils/base_argparse.py)."""
    psi, fixed = psi_at(i, anneal_psi, anneal_fn)
    if fixed:
        epoch = max(0, (i - anneal_psi) / period)
        return {"psi": psi, "stn_lr": decaying_cosine_lr(epoch, stn_lr, tm, decay), "ll_lr": decaying_cosine_lr(epoch, ll_lr, tm, decay)}
    return {"psi": psi, "stn_lr": stn_lr, "ll_lr": ll_lr}
