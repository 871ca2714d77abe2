"""Host-side schedule of the row-gather assembly pass (fem_shell_b200/csrc/fs_gather_plan.cpp) checked on the CPU
through the C-ABI hook fs_gather_plan: every (element, node row) incidence is scheduled exactly once, a warp's rows
fit its shared-memory slice, slots point at the right column nodes, and no two lanes of a warp add into the same
6x6 block in the same step of the same phase (the property that lets the kernel accumulate without atomics)."""
import numpy as np
import pytest

import fem_shell_b200 as fsb
import meshes

WARP_VALS = 2816


def structured(kind, nx, ny):
    m = fsb.meshgen(kind, nx, ny, 0, 0, 10, 7, (1, 0, 1, 0), 1.0, 2, 1)
    return m


MESHES = {
    "quad_plate": lambda: structured("q", 23, 17),
    "tri_plate": lambda: structured("t", 19, 14),
    "folded_mixed": lambda: meshes.folded_cantilever(skew=0.35),
    "umbrella_mixed": lambda: meshes.umbrella(mixed=True, n_rings=5),
    "umbrella_quads": lambda: meshes.umbrella(mixed=False, n_rings=4),
    "delaunay_mixed": lambda: meshes.delaunay_patch(400, seed=7, quad_fraction=0.3),
    "delaunay_tris": lambda: meshes.delaunay_patch(300, seed=11, quad_fraction=0.0),
}


def check_plan(m, plan):
    etype, eptr, enodes = m["etype"], m["eptr"], m["enodes"]
    n_nodes = m["xyz"].shape[0]
    nptr, nadj = plan["nptr"], plan["nadj"]
    info, nodes, chunks = plan["info"], plan["nodes"], plan["chunks"]
    # the pattern the slots refer to: sorted unique neighbours
    for p in range(n_nodes):
        row = nadj[nptr[p]:nptr[p + 1]]
        assert np.all(np.diff(row) > 0)
    expected = {}
    for e in range(etype.size):
        en = tuple(enodes[eptr[e]:eptr[e + 1]])
        for I in range(len(en)):
            expected[(en, I)] = expected.get((en, I), 0) + 1
    seen = {}
    row_cursor = 0
    for ci in range(chunks.shape[0]):
        val_off, val_count, n_phases, n_threads = chunks[ci]
        assert 0 < n_threads <= 32 and 0 < val_count <= WARP_VALS and val_count % 36 == 0
        assert val_off == 36 * nptr[row_cursor], "chunks cover the rows in order, contiguously"
        targets = {}
        rows_here = set()
        for lane in range(32):
            meta, rowinfo, mbits, slots = (int(v) for v in info[ci * 32 + lane])
            if lane >= n_threads:
                assert (meta >> 8) & 1 == 0
                continue
            assert (meta >> 8) & 1 == 1
            I, is_quad, phase = meta & 3, (meta >> 2) & 1, (meta >> 3) & 31
            nen = 4 if is_quad else 3
            assert phase < n_phases and I < nen
            en = tuple(int(v) for v in nodes[ci * 32 + lane][:nen])
            seen[(en, I)] = seen.get((en, I), 0) + 1
            row = en[I]
            rows_here.add(row)
            soff, deg = rowinfo & 0xffff, (rowinfo >> 16) & 0xffff
            assert deg == nptr[row + 1] - nptr[row]
            assert soff == 36 * (nptr[row] - nptr[row_cursor]) and soff + 36 * deg <= val_count
            for j in range(nen):
                slot = (slots >> (8 * j)) & 0xff
                assert slot < deg and nadj[nptr[row] + slot] == en[j], "slot j addresses the block of the element's j-th node"
                assert (mbits >> (8 * j)) & 0xff == 0
                key = (is_quad, phase, j, row, slot)
                assert key not in targets, "two lanes add into one block in the same step of the same phase"
                targets[key] = lane
        assert rows_here == set(range(row_cursor, row_cursor + len(rows_here))), "a chunk is a run of consecutive rows"
        assert sum(36 * (nptr[r + 1] - nptr[r]) for r in rows_here) == val_count
        row_cursor += len(rows_here)
    assert row_cursor == n_nodes
    assert seen == expected, "every incidence exactly once"


@pytest.mark.parametrize("name", sorted(MESHES))
def test_schedule_invariants(name):
    m = MESHES[name]()
    plan = fsb.gather_plan(m["etype"], m["eptr"], m["enodes"], m["xyz"].shape[0], warp_vals=WARP_VALS)
    assert plan is not None
    check_plan(m, plan)


def test_structured_quads_need_one_phase_and_fill_the_warp():
    m = structured("q", 40, 30)
    plan = fsb.gather_plan(m["etype"], m["eptr"], m["enodes"], m["xyz"].shape[0])
    assert plan["chunks"][:, 2].max() == 1            # the four elements around a node never meet in a slot at the same step
    interior = plan["chunks"][plan["chunks"][:, 1] == 8 * 9 * 36]
    assert interior.shape[0] > 0 and np.all(interior[:, 3] == 32)   # 8 interior rows x 4 incidences = a full warp


def test_meshgen_triangles_need_few_phases():
    m = structured("t", 30, 30)
    plan = fsb.gather_plan(m["etype"], m["eptr"], m["enodes"], m["xyz"].shape[0])
    assert 2 <= plan["chunks"][:, 2].max() <= 3


def test_hub_phases_and_fallbacks():
    m = meshes.umbrella(n_spokes=14, mixed=True, n_rings=5)
    plan = fsb.gather_plan(m["etype"], m["eptr"], m["enodes"], m["xyz"].shape[0])
    assert plan["chunks"][:, 2].max() == 14           # 14 triangles hold the hub as their first node: all clash in step 0
    # 40 elements at one node exceed a warp's 32 lanes: no schedule, the context falls back to the coloured pass
    big = meshes.umbrella(n_spokes=40, mixed=True, n_rings=3)
    assert fsb.gather_plan(big["etype"], big["eptr"], big["enodes"], big["xyz"].shape[0]) is None
    # a row that does not fit the shared-memory slice is refused as well
    assert fsb.gather_plan(m["etype"], m["eptr"], m["enodes"], m["xyz"].shape[0], warp_vals=36 * 10) is None


def test_dirichlet_bits_travel_with_the_entries():
    m = structured("q", 6, 5)
    n = m["xyz"].shape[0]
    mask = (np.arange(n) % 5 == 0).astype(np.uint8) * 0x3f | (np.arange(n) % 7 == 0).astype(np.uint8) * 0x07
    plan = fsb.gather_plan(m["etype"], m["eptr"], m["enodes"], n, mask=mask)
    info, nodes = plan["info"], plan["nodes"]
    for k in range(info.shape[0]):
        if (info[k, 0] >> 8) & 1:
            for j in range(4):
                assert (int(info[k, 2]) >> (8 * j)) & 0xff == mask[nodes[k, j]]


def test_plan_does_not_depend_on_the_number_of_host_threads(monkeypatch):
    """rows and chunks are planned on several host threads for large meshes; the schedule is a function of the mesh alone"""
    for m in (structured("t", 60, 45), meshes.umbrella(mixed=True, n_rings=5), structured("q", 70, 51)):
        plans = []
        for threads in ("1", "3", "7"):
            monkeypatch.setenv("FS_PLAN_THREADS", threads)
            plans.append(fsb.gather_plan(m["etype"], m["eptr"], m["enodes"], m["xyz"].shape[0]))
        for other in plans[1:]:
            for key in plans[0]:
                assert np.array_equal(plans[0][key], other[key]), key
    monkeypatch.setenv("FS_PLAN_THREADS", "5")
    m = MESHES["folded_mixed"]()
    check_plan(m, fsb.gather_plan(m["etype"], m["eptr"], m["enodes"], m["xyz"].shape[0]))
    big = meshes.umbrella(n_spokes=40, mixed=True, n_rings=3)
    assert fsb.gather_plan(big["etype"], big["eptr"], big["enodes"], big["xyz"].shape[0]) is None
