"""CPU test of the tile-binned grid ALGORITHM (what csrc/grid_binned.cuh computes), restated in numpy from the plan the
library reports (nvp_grid_bin_plan) and compared with the oracle's DenseGrid forward / backward.

It pins the index logic independently of the CUDA code: bucketing by tile, the per-level window with its aliased
"virtual" cells (res, j) / (i, res) (unclamped flat index, then modulo the level size - SURVEY A.2), the direct-access
branch for samples outside their tile's window (coordinates outside [0,1]) and the flush of a window onto the table.
The scatter-add's bank-interleaved layout with per-lane row ownership is restated lane by lane as well (one lane per
region address, flush through the table of populated entries).
The GPU tests check that the kernels implement exactly this (tests/test_gpu_binned.py)."""
import numpy as np
import pytest
import torch

from nvp_b200 import _lib
from oracle import nvp_oracle as O


def wrap(flat, cells):
    return np.mod(flat, cells)


def emulate(u, table, F, plan, params=None, dz=None):
    """One keyframe plane.  Either params [cells*F] -> gather output [n, L*F], or dz [n, L*F] -> gradient table [cells*F]
    (one direction per call: the window holds table values in the gather and partial sums in the scatter-add)."""
    assert (params is None) != (dz is None)
    tb, ext = plan["tiles_per_axis"], plan["window_extent"]
    n, L = u.shape[0], table.n_levels
    out = np.zeros((n, L * F), np.float64) if params is not None else None
    grad = np.zeros(int(table.offsets[-1]) * F, np.float64) if dz is not None else None
    b = np.clip((u * np.float32(tb)).astype(np.int32), 0, tb - 1)        # bin_tile_axis (NaN-free inputs)
    bucket = b[:, 1] * tb + b[:, 0]
    for tile in np.unique(bucket):
        sel = np.nonzero(bucket == tile)[0]
        ub = np.array([tile % tb, tile // tb], np.float32) / np.float32(tb)
        for l in range(L):
            s, res, E, off = table.scales[l], int(table.res[l]), ext[l], int(table.offsets[l])
            cells = res * res
            lo = np.floor(O._fmaf(np.full(2, s, np.float32), ub, 0.5)).astype(np.int64)
            amax = np.minimum(E - 2, res - 1 - lo)
            # window <-> table map, including the virtual cells; cells beyond (res, res) do not exist
            a_idx, b_idx = np.meshgrid(np.arange(E), np.arange(E), indexing="xy")
            g0, g1 = lo[0] + a_idx, lo[1] + b_idx
            ok = (g0 <= res) & (g1 <= res)
            glob = off + wrap(g0 + g1 * res, cells)
            win = np.zeros((E, E, F), np.float64)                          # [b, a, f]
            if params is not None:
                tabv = params.reshape(-1, F)
                win[ok] = tabv[glob[ok]]
            pos = O._fmaf(np.full((len(sel), 2), s, np.float32), u[sel], 0.5)
            fl = np.floor(pos)
            w = (pos - fl).astype(np.float32)
            i = fl.astype(np.int64)
            aa, bb = i[:, 0] - lo[0], i[:, 1] - lo[1]
            inside = (aa >= 0) & (aa <= amax[0]) & (bb >= 0) & (bb <= amax[1])
            for k, smp in enumerate(sel):
                for c1 in (0, 1):
                    for c0 in (0, 1):
                        wt = float((w[k, 0] if c0 else np.float32(1) - w[k, 0]) * (w[k, 1] if c1 else np.float32(1) - w[k, 1]))
                        if inside[k]:
                            cell = (bb[k] + c1, aa[k] + c0)
                            if params is not None:
                                out[smp, l * F:(l + 1) * F] += wt * win[cell]
                            if dz is not None:
                                win[cell] += wt * dz[smp, l * F:(l + 1) * F]
                        else:                                               # direct global access, modulo the level
                            gi = off + int(wrap((i[k, 0] + c0) + (i[k, 1] + c1) * res, cells))
                            if params is not None:
                                out[smp, l * F:(l + 1) * F] += wt * params.reshape(-1, F)[gi]
                            if dz is not None:
                                grad.reshape(-1, F)[gi] += wt * dz[smp, l * F:(l + 1) * F]
            if dz is not None:                                              # flush: one add per window cell
                np.add.at(grad.reshape(-1, F), glob[ok], win[ok])
    return out, grad


@pytest.mark.parametrize("F", [2, 4])
def test_binned_algorithm_equals_the_oracle_including_edges_and_out_of_range(F):
    L = 16
    d = _lib.NvpDesc(F, L, 16, 1.35, F, 6, 20, 24, 128, 3, 30.0)
    plan = _lib.grid_bin_plan(d, 4096)
    assert plan
    table = O.level_table(L, 16, 1.35)
    rng = np.random.default_rng(F)
    vals = np.array([0.0, 1.0, 0.5, 63.0 / 64, 1.0 / 1079, 1078.0 / 1079, 127.0 / 128], np.float32)
    edge = np.stack(np.meshgrid(vals, vals), -1).reshape(-1, 2)
    u = np.concatenate([rng.random((260, 2)).astype(np.float32), edge,
                        (rng.random((40, 2)) * 1.6 - 0.3).astype(np.float32)])          # some outside [0,1]
    n = u.shape[0]
    params = torch.from_numpy(rng.standard_normal(int(table.offsets[-1]) * F)).double().requires_grad_(True)
    dz = rng.standard_normal((n, L * F))
    ref = O.dense_grid_forward(params, torch.from_numpy(u), F, table)
    (ref * torch.from_numpy(dz)).sum().backward()
    out, _ = emulate(u, table, F, plan, params=params.detach().numpy())      # gather: windows loaded from the table
    _, grad = emulate(u, table, F, plan, dz=dz)                               # scatter-add: windows start at zero
    assert np.abs(out - ref.detach().numpy()).max() <= 1e-9
    assert np.abs(grad - params.grad.numpy()).max() <= 1e-9


def test_interleaved_scatter_layout_is_injective_and_bank_conflict_free():
    """The scatter-add's bank-interleaved window layout (grid_binned.cuh, ILV): cell (aa, row) of level l sits in slot
    (row >> 1) * E_l + aa of the bank pair (l & 7, row & 1) of its level group's region rows.  Restated here from the plan:
    distinct window cells never share an address, and the 16 lanes of a half-warp (8 levels x the two corner rows of one
    sample) always touch 16 different 8-byte bank pairs, whatever the sample's position in its windows."""
    L = 16
    d = _lib.NvpDesc(2, L, 16, 1.35, 2, 6, 20, 24, 128, 3, 30.0)
    ext = _lib.grid_bin_plan(d, 4096)["window_extent"]
    rows = [max((E + 1) // 2 * E for E in ext[:8]), max((E + 1) // 2 * E for E in ext[8:])]

    def address(l, aa, row):      # float offset inside the warp's region
        return (rows[0] * 32 if l >= 8 else 0) + ((l & 7) << 2) + ((((row >> 1) * ext[l] + aa) << 5) + ((row & 1) << 1))

    seen = set()
    for l in range(L):
        for row in range(ext[l]):
            for aa in range(ext[l]):
                a = address(l, aa, row)
                assert a not in seen and a + 1 not in seen and a % 2 == 0 and a + 1 < (rows[0] + rows[1]) * 32
                seen.update((a, a + 1))
    rng = np.random.default_rng(0)
    for _ in range(200):
        for half in (0, 1):
            banks = []
            for l in range(8 * half, 8 * half + 8):
                aa, bb = rng.integers(0, ext[l] - 1, 2)          # top-left corner; the lanes take rows bb and bb + 1
                for c1 in (0, 1):
                    banks.append((address(l, aa + rng.integers(0, 2), bb + c1) // 2) % 16)
            assert len(set(banks)) == 16


def emulate_interleaved_scatter(u, table, plan, dz):
    """The scatter-add of grid_binned_kernel<2, true, ..., ILV> for one plane, lane by lane: per task a zeroed region of
    128-byte rows, lane (level, parity) takes the window row of its own parity, adds into slot (row >> 1) * E + aa of
    its bank pair, samples outside the window go to the table directly, and the flush walks the table of populated
    (slot, pair) entries.  Returns the gradient table [cells * 2]."""
    F, L = 2, table.n_levels
    tb, ext = plan["tiles_per_axis"], plan["window_extent"]
    rows = [max((E + 1) // 2 * E for E in ext[:8]), max((E + 1) // 2 * E for E in ext[8:])]
    grad = np.zeros((int(table.offsets[-1]), F), np.float64)
    # flush table, built as the kernel builds it: row-major over (region row, bank pair)
    items = []
    for row in range(rows[0] + rows[1]):
        g, slot = (1, row - rows[0]) if row >= rows[0] else (0, row)
        for p in range(16):
            l = g * 8 + (p >> 1)
            if l < L:
                q, aa = divmod(slot, ext[l])
                wrow = 2 * q + (p & 1)
                if wrow < ext[l]:
                    items.append((row * 32 + p * 2, l, aa, wrow))
    assert len(items) == sum(E * E for E in ext)
    b = np.clip((u * np.float32(tb)).astype(np.int32), 0, tb - 1)
    bucket = b[:, 1] * tb + b[:, 0]
    for tile in np.unique(bucket):
        sel = np.nonzero(bucket == tile)[0]
        ub = np.array([tile % tb, tile // tb], np.float32) / np.float32(tb)
        region = np.zeros((rows[0] + rows[1]) * 32, np.float64)
        owner = {}                                                          # region address -> the only lane that may touch it
        lo = {}
        for l in range(L):
            lo[l] = np.floor(O._fmaf(np.full(2, table.scales[l], np.float32), ub, 0.5)).astype(np.int64)
        for smp in sel:
            for lane in range(32):
                l, par = lane >> 1, lane & 1
                s, res, E, off = table.scales[l], int(table.res[l]), ext[l], int(table.offsets[l])
                pos = O._fmaf(np.full(2, s, np.float32), u[smp], 0.5)
                fl = np.floor(pos)
                w = (pos - fl).astype(np.float32)
                i0, i1 = int(fl[0]), int(fl[1])
                aa, bb = i0 - lo[l][0], i1 - lo[l][1]
                c1 = (bb ^ par) & 1
                wr = w[1] if c1 else np.float32(1) - w[1]
                ka, kb = float((np.float32(1) - w[0]) * wr), float(w[0] * wr)
                d = dz[smp, l * F:(l + 1) * F]
                amax = np.minimum(E - 2, res - 1 - lo[l])
                if 0 <= aa <= amax[0] and 0 <= bb <= amax[1]:
                    wrow = bb + c1
                    assert (wrow & 1) == par
                    base = (rows[0] * 32 if lane >= 16 else 0) + ((l & 7) << 2) + (par << 1)
                    a0 = base + (((wrow >> 1) * E + aa) << 5)
                    for addr, k in ((a0, ka), (a0 + 32, kb)):
                        assert owner.setdefault(addr, lane) == lane
                        region[addr:addr + 2] += k * d
                else:                                                       # revisited by the direct path after the batch
                    cells = res * res
                    for c0, k in ((0, ka), (1, kb)):
                        grad[off + int(wrap(i0 + c0 + (i1 + c1) * res, cells))] += k * d
        for addr, l, aa, wrow in items:                                     # flush
            v = region[addr:addr + 2]
            res = int(table.res[l])
            g0, g1 = lo[l][0] + aa, lo[l][1] + wrow
            if (v != 0).any() and g0 <= res and g1 <= res:
                grad[int(table.offsets[l]) + int(wrap(g0 + g1 * res, res * res))] += v
    return grad.reshape(-1)


def test_interleaved_scatter_algorithm_equals_the_oracle():
    L, F = 16, 2
    d = _lib.NvpDesc(F, L, 16, 1.35, F, 6, 20, 24, 128, 3, 30.0)
    plan = _lib.grid_bin_plan(d, 4096)
    table = O.level_table(L, 16, 1.35)
    rng = np.random.default_rng(7)
    vals = np.array([0.0, 1.0, 0.5, 63.0 / 64, 1.0 / 1079, 1078.0 / 1079, 127.0 / 128], np.float32)
    edge = np.stack(np.meshgrid(vals, vals), -1).reshape(-1, 2)
    clustered = (np.float32(0.37) + rng.random((60, 2)) * np.float32(0.004)).astype(np.float32)   # many samples per window
    u = np.concatenate([rng.random((120, 2)).astype(np.float32), edge, clustered,
                        (rng.random((20, 2)) * 1.6 - 0.3).astype(np.float32)])
    params = torch.zeros(int(table.offsets[-1]) * F, dtype=torch.float64, requires_grad=True)
    dz = rng.standard_normal((u.shape[0], L * F))
    (O.dense_grid_forward(params, torch.from_numpy(u), F, table) * torch.from_numpy(dz)).sum().backward()
    grad = emulate_interleaved_scatter(u, table, plan, dz)
    assert np.abs(grad - params.grad.numpy()).max() <= 1e-9


def test_chained_scan_restated():
    """grid_bin_scan_kernel: 1024 buckets per CTA, CTAs chained through published (samples, tasks) totals in ticket order.
    Restated in numpy with the CTAs processed in an arbitrary ticket order: offsets, task list and totals equal the plain
    sequential scan's whatever order the tickets were handed out in."""
    rng = np.random.default_rng(3)
    m, chunk = 3 * 32 * 32, 160
    cnt = rng.integers(0, 400, m)
    cnt[rng.random(m) < 0.3] = 0
    tasks_of = lambda c: 0 if c == 0 else (1 if c <= chunk else -(-c // chunk))
    # sequential reference
    offs_ref = np.concatenate([[0], np.cumsum(cnt)])
    tasks_ref = [(i, offs_ref[i] + b) for i in range(m) for b in range(0, cnt[i], chunk)]
    # chained version: CTA with ticket b handles buckets [1024 b, 1024 (b + 1)) -- the ticket IS the logical block index
    nblk = m // 1024
    published = {}
    offs = np.zeros(m + 1, np.int64)
    tasks = {}
    for b in range(nblk):                                   # tickets are taken in scheduling order
        c = cnt[b * 1024:(b + 1) * 1024]
        t = np.array([tasks_of(x) for x in c])
        published[b] = (int(c.sum()), int(t.sum()))
        cpre = sum(published[k][0] for k in range(b))       # every predecessor has a lower ticket: already published
        tpre = sum(published[k][1] for k in range(b))
        off = cpre + np.cumsum(c) - c
        to = tpre + np.cumsum(t) - t
        for j in range(1024):
            i = b * 1024 + j
            offs[i] = off[j]
            for k, bb in enumerate(range(0, c[j], chunk)):
                tasks[to[j] + k] = (i, off[j] + bb)
        if b == nblk - 1:
            offs[m] = off[-1] + c[-1]
            n_tasks = to[-1] + t[-1]
    assert (offs == offs_ref).all()
    assert n_tasks == len(tasks_ref) and [tasks[k] for k in range(n_tasks)] == tasks_ref
