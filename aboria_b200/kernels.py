"""Kernel functions F(dx, a, b) of sparse operators, by name.  Each maps to a
device functor in include/aboria_b200/device_kernel.cuh (the reference takes a
host lambda, src/Operators.h:478-516).  Variable names refer to columns of the
row / column Particles; they are resolved when the operator is applied, so an
operator sees later reorders exactly like the reference's (which stores
references to the particle sets, src/Kernels.h:133-134)."""

K_CONST_SUM, K_CONST_SUM_DIFF, K_INV_DIST, K_INV_DIST_AA = 0, 1, 2, 3
K_WENDLAND_C2, K_LJ_FORCE, K_SPH_DENSITY, K_SPH_PRESSURE = 4, 5, 6, 7
K_LINEAR_SPRING = 8


class Kernel:
    def __init__(self, kernel_id, block_rows, block_cols, params=(), row_vars=(), col_vars=()):
        self.kernel_id = kernel_id
        self.block_rows = block_rows
        self.block_cols = block_cols
        self.params = tuple(float(p) for p in params)
        self.row_vars = tuple(row_vars)
        self.col_vars = tuple(col_vars)


def const_sum(s1, s2):
    """get<s1>(a) + get<s2>(b)  (tests/operators.h:842-847)"""
    return Kernel(K_CONST_SUM, 1, 1, (), (s1,), (s2,))


def const_sum_diff(s1, s2):
    """2x1 block (s1(a)+s2(b), s1(a)-s2(b))  (tests/operators.h:905-911)"""
    return Kernel(K_CONST_SUM_DIFF, 2, 1, (), (s1,), (s2,))


def inv_dist(eps):
    """1/(|dx| + eps)"""
    return Kernel(K_INV_DIST, 1, 1, (eps,))


def inv_dist_aa(eps, a):
    """a_i a_j/(|dx| + eps)  (tests/operators.h:251-256)"""
    return Kernel(K_INV_DIST_AA, 1, 1, (eps,), (a,), (a,))


def wendland_c2(h):
    """(2-|dx|/h)^4 (1+2|dx|/h)  (tests/rbf_interpolation.h:310-313)"""
    return Kernel(K_WENDLAND_C2, 1, 1, (h,))


def lj_force(D, sigma, eps):
    """D x 1 Lennard-Jones force block (tests/md.h:166-174 pattern)"""
    return Kernel(K_LJ_FORCE, D, 1, (sigma, eps))


def sph_density(h, mass, wcon):
    """mass * W(|dx|, h)  (tests/sph.h:154-165)"""
    return Kernel(K_SPH_DENSITY, 1, 1, (h, mass, wcon))


def sph_pressure(D, h, mass, wcon, pdr2):
    """D x 1: mass (pdr2_a + pdr2_b) F(|dx|,h) dx  (tests/sph.h:140-152, :333-339)"""
    return Kernel(K_SPH_PRESSURE, D, 1, (h, mass, wcon), (pdr2,), (pdr2,))


def linear_spring(D, k, diameter):
    """D x 1 linear spring force -k (diameter/|dx| - 1) dx, 0 at |dx| = 0 (tests/md.h:166-174)"""
    return Kernel(K_LINEAR_SPRING, D, 1, (k, diameter))
