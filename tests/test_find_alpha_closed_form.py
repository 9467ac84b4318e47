"""find_alpha (low_order_graph_viscosity.jl:299-327): the closed form the CUDA kernels use (csrc/kernels2d.cuh:
find_alpha_closed) against the reference's doubling + 50-step bisection, both restated in numpy, on random state pairs.
The bisection's predicate is evaluated on differences of size POSTOL, so the two agree to its own evaluation noise."""
import numpy as np

POSTOL = 1e-14


def rhoe(s):
    return s[3] - 0.5 * (s[1] ** 2 + s[2] ** 2) / s[0]


def bisect(ui, ut):
    aL, aR = 0.0, 1.0
    ok = lambda a: (a * ui - ut)[0] > POSTOL and rhoe(a * ui - ut) > POSTOL
    while not ok(aR):
        aR *= 2
    for _ in range(50):
        aM = (aL + aR) / 2
        if ok(aM):
            aR = aM
        else:
            aL = aM
    return aR


def closed(ui, ut, eps=POSTOL):
    d = ut - ui
    rho, E = ui[0], ui[3]
    msq, mdm, dmsq = ui[1] ** 2 + ui[2] ** 2, ui[1] * d[1] + ui[2] * d[2], d[1] ** 2 + d[2] ** 2
    beta = (d[0] + eps) / rho
    A = E * rho - 0.5 * msq
    B = -(E * d[0] + d[3] * rho) + mdm - eps * rho
    Cq = d[3] * d[0] - 0.5 * dmsq + eps * d[0]
    disc = B * B - 4 * A * Cq
    if disc >= 0:
        sq = np.sqrt(disc)
        beta = max(beta, (sq - B) / (2 * A) if B <= 0 else -2 * Cq / (B + sq))
    return max(1 + beta, 2.0 ** -50)


def cons(rho, u, v, p, g=1.4):
    return np.array([rho, rho * u, rho * v, p / (g - 1) + 0.5 * rho * (u * u + v * v)])


def test_closed_form_matches_bisection():
    rng = np.random.default_rng(7)
    worst = 0.0
    for _ in range(3000):
        ui = cons(rng.uniform(0.1, 8), rng.uniform(-8, 8), rng.uniform(-8, 8), rng.uniform(0.05, 100))
        scale = 10.0 ** rng.uniform(-12, -0.3)         # u~ from 1e-12 to 50 % away from u, as the projected face values are
        ut = ui + scale * np.abs(ui).max() * rng.standard_normal(4) * np.array([1, 1, 1, 1.0])
        if ut[0] <= 0 or rhoe(ut) <= 0:
            continue
        a, b = bisect(ui, ut), closed(ui, ut)
        worst = max(worst, abs(a - b) / a)
    assert worst < 1e-12, worst


def test_identical_states_give_one_plus_eps():
    ui = cons(1.3, 0.4, -0.2, 0.9)
    assert abs(closed(ui, ui.copy()) - bisect(ui, ui.copy())) < 1e-13
