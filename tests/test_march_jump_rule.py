"""CPU: the arithmetic of the NeuS march's jump rules (cnrma_stage_b.cu, march_neus_kernel<SKIP, FINE>) restated in numpy
float32 and checked against the property that makes a jump exact: every sample that is jumped over reads the same
sigmoid(-tsdf) value as the sample the jump started from (so alpha == 0 there, rm.py:757-762), and lies inside the grid.

The clearance field used here is the LARGEST one the definition allows (brute force), so the jumps are at least as long
as the kernel's, whose separable distance transform can only be more conservative."""
import numpy as np
import pytest

F = np.float32
MARGIN = F(0.05)   # kSkipMargin
CAP = 15           # kDistCap


def _clearance(s):
    """D(c) = largest k <= CAP such that every voxel within L-infinity distance k of c is inside the grid and holds s(c)."""
    nx, ny, nz = s.shape
    D = np.zeros(s.shape, np.int32)
    for x in range(nx):
        for y in range(ny):
            for z in range(nz):
                k = 0
                while k < CAP:
                    r = k + 1
                    if x - r < 0 or y - r < 0 or z - r < 0 or x + r >= nx or y + r >= ny or z + r >= nz:
                        break
                    if not np.all(s[x - r:x + r + 1, y - r:y + r + 1, z - r:z + r + 1] == s[x, y, z]):
                        break
                    k = r
                D[x, y, z] = k
    return D


def _blocky_tsdf(rng, dim):
    """Free space / unobserved space at +-0.999 with a few slabs and boxes of other values, like the head's output."""
    t = np.full(dim, 0.999, np.float32)
    for _ in range(4):
        lo = [int(rng.integers(0, d - 2)) for d in dim]
        hi = [int(min(d, l + rng.integers(1, max(2, d // 2)))) for d, l in zip(dim, lo)]
        t[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = F(rng.choice([-0.999, -0.4, 0.0, 0.3, 0.999]))
    # a thin graded band (every voxel differs from its neighbours): clearance 0, the k == 0 case of the FINE rule
    z0 = int(rng.integers(1, dim[2] - 1))
    t[:, :, z0] = np.linspace(-0.9, 0.9, dim[0] * dim[1], dtype=np.float32).reshape(dim[0], dim[1])
    return t


def _sample_voxel(o, d, t, origin, vs):
    """rm.py:729-733 in float32: position, (p - origin) / vs, round half to even; also the kernel's q = rel * RN(1 / vs)."""
    p = (o + d * t).astype(np.float32)              # float32 multiply, then float32 add (no contraction)
    rel = (p - origin).astype(np.float32)
    r = np.rint((rel / vs).astype(np.float32))
    q = (rel * (F(1.0) / vs)).astype(np.float32)
    return r.astype(np.int64), (q - r.astype(np.float32)).astype(np.float32)


def _jump_fine(frac, k, inv_step):
    room = F(k) + (F(0.5) - MARGIN)
    terms = [F(np.float64(-frac[a]) * np.float64(inv_step[a]) + np.float64(F(room * np.abs(inv_step[a])))) for a in range(3)]
    return int(np.trunc(min(terms)))


def _jump_coarse(k, inv_step_max):
    return int(np.trunc((F(k) - MARGIN) * inv_step_max)) if k > 0 else 0


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("fine", [False, True])
def test_jumped_samples_read_the_same_value(seed, fine):
    rng = np.random.default_rng(900 + seed)
    dim = (int(rng.integers(10, 22)), int(rng.integers(10, 22)), int(rng.integers(6, 14)))
    vs = F(rng.choice([0.04, 0.08, 0.16]))
    origin = (np.zeros(3) if seed % 2 == 0 else rng.uniform(-1, 1, 3)).astype(np.float32)
    tsdf = _blocky_tsdf(rng, dim)
    s = (F(1.0) / (F(1.0) + np.exp(tsdf))).astype(np.float32)
    D = _clearance(s)
    N = int(rng.choice([40, 97, 300, 900]))
    t_one = F(np.sqrt(float(dim[0] ** 2 + dim[1] ** 2 + dim[2] ** 2)) * float(vs) / N)
    extent = np.array(dim, np.float32) * vs
    jumps = checked = within_voxel = 0
    for _ in range(60):
        o = (origin + rng.uniform(-0.2, 1.2, 3).astype(np.float32) * extent).astype(np.float32)
        d = rng.normal(size=3).astype(np.float32)
        if rng.random() < 0.3:
            d[int(rng.integers(0, 3))] = 0.0            # axis-parallel in one coordinate
        d = (d / np.linalg.norm(d)).astype(np.float32)
        step = (d * t_one / vs).astype(np.float32)
        inv_step = np.copysign(F(1.0) / np.maximum(np.abs(step), F(1e-6)), step).astype(np.float32)
        step_max = np.abs(step).max()
        inv_step_max = F(1.0) / step_max if step_max > 0 else F(0.0)
        i = 0
        while i < N:
            v, frac = _sample_voxel(o, d, F(i) * t_one, origin, vs)
            inside = bool(np.all(v >= 0) and np.all(v < dim))
            n = 0
            if inside:
                k = int(D[tuple(v)])
                n = _jump_fine(frac, k, inv_step) if fine else _jump_coarse(k, inv_step_max)
                n = max(0, min(n, N - 1 - i))
                for j in range(1, n + 1):
                    vj, _ = _sample_voxel(o, d, F(i + j) * t_one, origin, vs)
                    assert np.all(vj >= 0) and np.all(vj < dim), (i, j, v, vj, k)
                    assert s[tuple(vj)] == s[tuple(v)], (i, j, v, vj, k)
                    checked += 1
                    within_voxel += int(np.array_equal(vj, v))
                jumps += n > 0
            i += n + 1
    assert jumps > 0 and checked > 0
    if fine and N >= 300:
        assert within_voxel > 0     # the k == 0 / same-voxel jumps of the FINE rule are exercised


def test_fine_rule_never_jumps_less_than_a_voxel_allows():
    """With the sample at the voxel centre and an axis-parallel ray, the FINE rule gives floor((k + 0.45) / step) samples --
    at least the coarse rule's floor((k - 0.05) / step)."""
    for k in range(0, 16):
        for step in (0.13, 0.39, 0.77, 1.25):
            inv = np.array([F(1.0) / F(step), F(1e6), F(1e6)], np.float32)
            fine = _jump_fine(np.zeros(3, np.float32), k, inv)
            coarse = _jump_coarse(k, F(1.0) / F(step))
            assert fine >= coarse
            assert fine == int(np.floor((k + 0.45) / step + 1e-4)) or fine == int(np.floor((k + 0.45) / step - 1e-4))
