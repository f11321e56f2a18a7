"""Oracle restatement of the cell-grouping extraction (oracle/port/port.cpp, the maps of the second
extract_iso_mesh / extract_MI_mesh overloads, src/extract_mesh.cpp:268-566 and :988-1443) pinned against
digests of the reference's own functions (tests/golden/cellgroup_golden.json, made by
tests/golden/make_cellgroup_golden.py) and, where the hybrid reference library is present, element-wise."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from helpers import MAP_I64, make_funcs, orc_eval, orc_grid, orc_run, ref_lib  # noqa: E402
from make_cellgroup_golden import CASES, reference_maps, sha  # noqa: E402

with open(os.path.join(HERE, "golden", "cellgroup_golden.json")) as _f:
    GOLD = json.load(_f)


@pytest.mark.parametrize("name", sorted(CASES))
def test_port_maps_match_the_reference_digests(name):
    mode, R, specs = CASES[name]
    pts, tets = orc_grid(R)
    vals = orc_eval(make_funcs(specs), pts)
    port = orc_run(mode, pts, tets, vals)
    assert port.error == ""
    assert port["stats"][-2:].tolist() == GOLD[name]["counts"]
    for k in MAP_I64:
        assert sha(port[k]) == GOLD[name][k], k
    # every local vertex of an active tet has an entry, the first four are the tet corners
    start = port["global_vId_start_index_of_tet"]
    sizes = np.diff(start)
    active = np.nonzero(sizes)[0]
    assert (sizes[active] >= 4).all() and len(start) == len(tets) + 1
    corners = port["global_vId_of_tet_vert"][(start[active][:, None] + np.arange(4)).ravel()].reshape(-1, 4)
    assert np.array_equal(-corners - 1, tets[active].astype(np.int64))


@pytest.mark.skipif(ref_lib() is None, reason="hybrid reference not built")
@pytest.mark.parametrize("name", ["ia_plane_through_vertices_R8", "mi_ties_on_tet_faces_R8"])
def test_port_maps_equal_the_reference_function(name):
    mode, R, specs = CASES[name]
    pts, tets = orc_grid(R)
    vals = orc_eval(make_funcs(specs), pts)
    port = orc_run(mode, pts, tets, vals)
    ref = reference_maps(mode, R, specs)
    for k in MAP_I64:
        assert np.array_equal(port[k], ref[k]), k


@pytest.mark.skipif(ref_lib() is None, reason="hybrid reference not built")
@pytest.mark.parametrize("name", sorted(CASES))
def test_product_cell_graph_equals_the_reference_functions(name):
    """host/cell_graph.h (simplicial-cell adjacency + connected components) against the reference's
    build_simplicial_cell_adjacency / compute_simplicial_cell_connected_components (src/cell_connectivity.cpp:15-372)
    on the same complexes and maps: adjacency arrays element-wise, components as shell sets."""
    mode, R, specs = CASES[name]
    ref = reference_maps(mode, R, specs)
    assert ref.error == ""
    adjacency_equal, components_equal, n_cells, n_components = ref["cell_graph"].tolist()
    assert adjacency_equal == 1 and components_equal == 1
    assert n_cells >= 5 * R ** 3 and 1 <= n_components < n_cells
