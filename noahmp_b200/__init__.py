"""noahmp_b200 — B200-native (sm_100a CUDA) Noah-MP column physics behind the reference's `noahmplsm` interface.

Host-side mirror of the reference interface for this path (phys/module_sf_noahmpdrv.F90):
    NoahMP.noahmplsm(arrays, scalars)   <->  CALL noahmplsm(...)           (:11-844)
    read_tables(dir, dataset, soil)     <->  read_mp_veg_parameters + SOIL_VEG_GEN_PARM
    proc_grid / tile                    <->  mpp_land_get_nprocsxy / mpp_land_partition_calc
"""
from .driver import bind_numa, NoahMP, NoahMPDomain, NoahmpError, read_tables, proc_grid, tile, tile_neighbours, SYNC_FULL, SYNC_RESIDENT, MATH_FAST, MATH_PARITY, HINT_DZ8W_CONSTANT, HINT_VEGFRA_UNCHANGED, HINT_P8W_LEVELS_EQUAL  # noqa: F401
