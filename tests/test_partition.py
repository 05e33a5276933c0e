"""Domain decomposition = mpp/module_mpp_land.F90 (mpp_land_get_nprocsxy :124-141, mpp_land_partition_calc
:227-288): integer maps, bit-exact.  The expected values are the restated arithmetic of SURVEY.md §4/§8e (the
reference's own test/test_mpp_land_partition.F90 prints these for a 101x101 domain and asserts nothing)."""
import numpy as np
import pytest

import noahmp_b200


def ref_proc_grid(nproc):
    best, nx, ny = nproc, None, None
    for j in range(1, nproc + 1):
        if nproc % j == 0:
            i = nproc // j
            if abs(i - j) < best:
                best, nx, ny = abs(i - j), i, j
    return nx, ny


def ref_tiles(gnx, gny, nproc):
    npx, npy = ref_proc_grid(nproc)
    out = []
    for r in range(nproc):
        ipx, ipy = r % npx, r // npx
        nxs = [gnx // npx + (1 if k <= gnx % npx - 1 else 0) for k in range(npx)]
        nys = [gny // npy + (1 if k <= gny % npy - 1 else 0) for k in range(npy)]
        xs, ys = 1 + sum(nxs[:ipx]), 1 + sum(nys[:ipy])
        out.append((xs, xs + nxs[ipx] - 1, ys, ys + nys[ipy] - 1))
    return out


def test_process_grid_known_answers(built):
    want = {1: (1, 1), 2: (2, 1), 3: (3, 1), 4: (2, 2), 6: (3, 2), 8: (4, 2), 12: (4, 3), 16: (4, 4), 7: (7, 1)}
    for n, g in want.items():
        assert noahmp_b200.proc_grid(n) == g == ref_proc_grid(n)


def test_partition_101x101_known_answers(built):
    """test/test_mpp_land_partition.F90 domain: tiles {51,50} x 101 on 2 ranks, {51,50}^2 on 4, {26,25,25,25} x {51,50} on 8."""
    assert [noahmp_b200.tile(101, 101, 2, r) for r in range(2)] == [(1, 51, 1, 101), (52, 101, 1, 101)]
    assert [noahmp_b200.tile(101, 101, 4, r) for r in range(4)] == [(1, 51, 1, 51), (52, 101, 1, 51), (1, 51, 52, 101),
                                                                      (52, 101, 52, 101)]
    t8 = [noahmp_b200.tile(101, 101, 8, r) for r in range(8)]
    assert t8[0] == (1, 26, 1, 51) and t8[3] == (77, 101, 1, 51) and t8[4] == (1, 26, 52, 101) and t8[7] == (77, 101, 52, 101)


def test_conus_and_nldas_tiles(built):
    assert noahmp_b200.tile(4608, 3840, 2, 1) == (2305, 4608, 1, 3840)
    assert noahmp_b200.tile(4608, 3840, 4, 3) == (2305, 4608, 1921, 3840)
    assert noahmp_b200.tile(4608, 3840, 8, 5) == (1153, 2304, 1921, 3840)
    assert noahmp_b200.tile(464, 224, 8, 7) == (349, 464, 113, 224)


@pytest.mark.parametrize("gnx,gny", [(101, 101), (464, 224), (4608, 3840), (7, 5), (13, 29)])
@pytest.mark.parametrize("nproc", [1, 2, 3, 4, 5, 6, 8, 12, 16])
def test_tiles_cover_domain_exactly_once(built, gnx, gny, nproc):
    if nproc > min(gnx, gny):
        pytest.skip("more ranks than rows")
    tiles = [noahmp_b200.tile(gnx, gny, nproc, r) for r in range(nproc)]
    assert tiles == ref_tiles(gnx, gny, nproc)
    cover = np.zeros((gny, gnx), np.int32)
    for xs, xe, ys, ye in tiles:
        cover[ys - 1:ye, xs - 1:xe] += 1
    assert (cover == 1).all()
