!> module_sf_noahmpdrv_b200.F90 -- drop-in replacement of the `noahmplsm` grid loop of
!> phys/module_sf_noahmpdrv.F90:11-844 that forwards to libnoahmp_b200.so (include/noahmp_b200.h).
!>
!> The SUBROUTINE keeps the reference's name, dummy-argument list and order, so
!> driver/module_hrldas_noahmp_driver.F90:386-415 compiles and calls it unchanged.  Build: rename the
!> reference's `noahmplsm` (or drop its body), add this file to phys/Makefile, link with
!> -lnoahmp_b200 -lcudart -lstdc++.  NOT compiled in the build image of this repo (no Fortran compiler there).
!> Member and argument order follow noahmp_b200/_capi.py:ARGS_SPEC (= the struct in the header).
MODULE module_sf_noahmpdrv_b200
  USE, INTRINSIC :: ISO_C_BINDING
  IMPLICIT NONE
  PRIVATE
  PUBLIC :: noahmplsm, NOAHMP_INIT, WTABLE_mmf_noahmp
  PUBLIC :: noahmp_b200_start, noahmp_b200_start_parallel, noahmp_b200_stop, noahmp_b200_refresh_host
  PUBLIC :: noahmp_b200_snapshot_begin, noahmp_b200_snapshot_wait, noahmp_b200_vegfra_changed
  PUBLIC :: noahmp_b200_global_budget
  PUBLIC :: noahmp_forcing_fields, noahmp_b200_forcing_static, noahmp_b200_forcing_upload, noahmp_b200_forcing_swap
  PUBLIC :: noahmp_b200_forcing_apply, noahmplsm_device_forcing

  !> mirrors `noahmp_lsm_args` member for member
  TYPE, BIND(C) :: noahmp_lsm_args
    INTEGER(C_INT) :: itimestep
    INTEGER(C_INT) :: yr
    REAL(C_FLOAT) :: julian
    TYPE(C_PTR) :: coszin
    TYPE(C_PTR) :: xlatin
    TYPE(C_PTR) :: dz8w
    REAL(C_FLOAT) :: dt
    TYPE(C_PTR) :: dzs
    INTEGER(C_INT) :: nsoil
    REAL(C_FLOAT) :: dx
    TYPE(C_PTR) :: ivgtyp
    TYPE(C_PTR) :: isltyp
    TYPE(C_PTR) :: vegfra
    TYPE(C_PTR) :: vegmax
    TYPE(C_PTR) :: tmn
    TYPE(C_PTR) :: xland
    TYPE(C_PTR) :: xice
    REAL(C_FLOAT) :: xice_thres
    INTEGER(C_INT) :: isice
    INTEGER(C_INT) :: isurban
    INTEGER(C_INT) :: idveg
    INTEGER(C_INT) :: iopt_crs
    INTEGER(C_INT) :: iopt_btr
    INTEGER(C_INT) :: iopt_run
    INTEGER(C_INT) :: iopt_sfc
    INTEGER(C_INT) :: iopt_frz
    INTEGER(C_INT) :: iopt_inf
    INTEGER(C_INT) :: iopt_rad
    INTEGER(C_INT) :: iopt_alb
    INTEGER(C_INT) :: iopt_snf
    INTEGER(C_INT) :: iopt_tbot
    INTEGER(C_INT) :: iopt_stc
    INTEGER(C_INT) :: iz0tlnd
    TYPE(C_PTR) :: t3d
    TYPE(C_PTR) :: qv3d
    TYPE(C_PTR) :: u_phy
    TYPE(C_PTR) :: v_phy
    TYPE(C_PTR) :: swdown
    TYPE(C_PTR) :: glw
    TYPE(C_PTR) :: p8w3d
    TYPE(C_PTR) :: rainbl
    TYPE(C_PTR) :: tsk
    TYPE(C_PTR) :: hfx
    TYPE(C_PTR) :: qfx
    TYPE(C_PTR) :: lh
    TYPE(C_PTR) :: grdflx
    TYPE(C_PTR) :: smstav
    TYPE(C_PTR) :: smstot
    TYPE(C_PTR) :: sfcrunoff
    TYPE(C_PTR) :: udrunoff
    TYPE(C_PTR) :: albedo
    TYPE(C_PTR) :: snowc
    TYPE(C_PTR) :: smois
    TYPE(C_PTR) :: sh2o
    TYPE(C_PTR) :: tslb
    TYPE(C_PTR) :: snow
    TYPE(C_PTR) :: snowh
    TYPE(C_PTR) :: canwat
    TYPE(C_PTR) :: acsnom
    TYPE(C_PTR) :: acsnow
    TYPE(C_PTR) :: emiss
    TYPE(C_PTR) :: qsfc
    TYPE(C_PTR) :: isnowxy
    TYPE(C_PTR) :: tvxy
    TYPE(C_PTR) :: tgxy
    TYPE(C_PTR) :: canicexy
    TYPE(C_PTR) :: canliqxy
    TYPE(C_PTR) :: eahxy
    TYPE(C_PTR) :: tahxy
    TYPE(C_PTR) :: cmxy
    TYPE(C_PTR) :: chxy
    TYPE(C_PTR) :: fwetxy
    TYPE(C_PTR) :: sneqvoxy
    TYPE(C_PTR) :: alboldxy
    TYPE(C_PTR) :: qsnowxy
    TYPE(C_PTR) :: wslakexy
    TYPE(C_PTR) :: zwtxy
    TYPE(C_PTR) :: waxy
    TYPE(C_PTR) :: wtxy
    TYPE(C_PTR) :: tsnoxy
    TYPE(C_PTR) :: zsnsoxy
    TYPE(C_PTR) :: snicexy
    TYPE(C_PTR) :: snliqxy
    TYPE(C_PTR) :: lfmassxy
    TYPE(C_PTR) :: rtmassxy
    TYPE(C_PTR) :: stmassxy
    TYPE(C_PTR) :: woodxy
    TYPE(C_PTR) :: stblcpxy
    TYPE(C_PTR) :: fastcpxy
    TYPE(C_PTR) :: xlaixy
    TYPE(C_PTR) :: xsaixy
    TYPE(C_PTR) :: taussxy
    TYPE(C_PTR) :: smoiseq
    TYPE(C_PTR) :: smcwtdxy
    TYPE(C_PTR) :: deeprechxy
    TYPE(C_PTR) :: rechxy
    TYPE(C_PTR) :: t2mvxy
    TYPE(C_PTR) :: t2mbxy
    TYPE(C_PTR) :: q2mvxy
    TYPE(C_PTR) :: q2mbxy
    TYPE(C_PTR) :: tradxy
    TYPE(C_PTR) :: neexy
    TYPE(C_PTR) :: gppxy
    TYPE(C_PTR) :: nppxy
    TYPE(C_PTR) :: fvegxy
    TYPE(C_PTR) :: runsfxy
    TYPE(C_PTR) :: runsbxy
    TYPE(C_PTR) :: ecanxy
    TYPE(C_PTR) :: edirxy
    TYPE(C_PTR) :: etranxy
    TYPE(C_PTR) :: fsaxy
    TYPE(C_PTR) :: firaxy
    TYPE(C_PTR) :: aparxy
    TYPE(C_PTR) :: psnxy
    TYPE(C_PTR) :: savxy
    TYPE(C_PTR) :: sagxy
    TYPE(C_PTR) :: rssunxy
    TYPE(C_PTR) :: rsshaxy
    TYPE(C_PTR) :: bgapxy
    TYPE(C_PTR) :: wgapxy
    TYPE(C_PTR) :: tgvxy
    TYPE(C_PTR) :: tgbxy
    TYPE(C_PTR) :: chvxy
    TYPE(C_PTR) :: chbxy
    TYPE(C_PTR) :: shgxy
    TYPE(C_PTR) :: shcxy
    TYPE(C_PTR) :: shbxy
    TYPE(C_PTR) :: evgxy
    TYPE(C_PTR) :: evbxy
    TYPE(C_PTR) :: ghvxy
    TYPE(C_PTR) :: ghbxy
    TYPE(C_PTR) :: irgxy
    TYPE(C_PTR) :: ircxy
    TYPE(C_PTR) :: irbxy
    TYPE(C_PTR) :: trxy
    TYPE(C_PTR) :: evcxy
    TYPE(C_PTR) :: chleafxy
    TYPE(C_PTR) :: chucxy
    TYPE(C_PTR) :: chv2xy
    TYPE(C_PTR) :: chb2xy
    INTEGER(C_INT) :: ids
    INTEGER(C_INT) :: ide
    INTEGER(C_INT) :: jds
    INTEGER(C_INT) :: jde
    INTEGER(C_INT) :: kds
    INTEGER(C_INT) :: kde
    INTEGER(C_INT) :: ims
    INTEGER(C_INT) :: ime
    INTEGER(C_INT) :: jms
    INTEGER(C_INT) :: jme
    INTEGER(C_INT) :: kms
    INTEGER(C_INT) :: kme
    INTEGER(C_INT) :: its
    INTEGER(C_INT) :: ite
    INTEGER(C_INT) :: jts
    INTEGER(C_INT) :: jte
    INTEGER(C_INT) :: kts
    INTEGER(C_INT) :: kte
  END TYPE noahmp_lsm_args

  TYPE, BIND(C) :: noahmp_status
    INTEGER(C_INT) :: code, i, j, count
    REAL(C_FLOAT)  :: value
  END TYPE noahmp_status

  !> member for member include/noahmp_b200.h :: noahmp_wtable_args (the WTABLE_mmf_noahmp dummy list,
  !> phys/module_sf_noahmp_groundwater.F90:14-22)
  TYPE, BIND(C) :: noahmp_wtable_args
    INTEGER(C_INT) :: nsoil
    TYPE(C_PTR) :: xland
    TYPE(C_PTR) :: xice
    REAL(C_FLOAT) :: xice_threshold
    INTEGER(C_INT) :: isice
    TYPE(C_PTR) :: isltyp
    TYPE(C_PTR) :: smoiseq
    TYPE(C_PTR) :: dzs
    REAL(C_FLOAT) :: wtddt
    TYPE(C_PTR) :: fdepth
    TYPE(C_PTR) :: area
    TYPE(C_PTR) :: topo
    INTEGER(C_INT) :: isurban
    TYPE(C_PTR) :: ivgtyp
    TYPE(C_PTR) :: rivercond
    TYPE(C_PTR) :: riverbed
    TYPE(C_PTR) :: eqwtd
    TYPE(C_PTR) :: pexp
    TYPE(C_PTR) :: smois
    TYPE(C_PTR) :: sh2oxy
    TYPE(C_PTR) :: smcwtd
    TYPE(C_PTR) :: wtd
    TYPE(C_PTR) :: qrf
    TYPE(C_PTR) :: deeprech
    TYPE(C_PTR) :: qspring
    TYPE(C_PTR) :: qslat
    TYPE(C_PTR) :: qrfs
    TYPE(C_PTR) :: qsprings
    TYPE(C_PTR) :: rech
    INTEGER(C_INT) :: ids
    INTEGER(C_INT) :: ide
    INTEGER(C_INT) :: jds
    INTEGER(C_INT) :: jde
    INTEGER(C_INT) :: kds
    INTEGER(C_INT) :: kde
    INTEGER(C_INT) :: ims
    INTEGER(C_INT) :: ime
    INTEGER(C_INT) :: jms
    INTEGER(C_INT) :: jme
    INTEGER(C_INT) :: kms
    INTEGER(C_INT) :: kme
    INTEGER(C_INT) :: its
    INTEGER(C_INT) :: ite
    INTEGER(C_INT) :: jts
    INTEGER(C_INT) :: jte
    INTEGER(C_INT) :: kts
    INTEGER(C_INT) :: kte
  END TYPE noahmp_wtable_args

  !> one forcing file for the on-device forcing pipeline (include/noahmp_b200.h :: noahmp_forcing_fields)
  TYPE, BIND(C) :: noahmp_forcing_fields
    TYPE(C_PTR) :: t
    TYPE(C_PTR) :: q
    TYPE(C_PTR) :: u
    TYPE(C_PTR) :: v
    TYPE(C_PTR) :: p
    TYPE(C_PTR) :: lw
    TYPE(C_PTR) :: sw
    TYPE(C_PTR) :: pcp
    TYPE(C_PTR) :: fpar
  END TYPE noahmp_forcing_fields

  !> member for member include/noahmp_b200.h :: noahmp_init_args (the NOAHMP_INIT dummy list)
  TYPE, BIND(C) :: noahmp_init_args
    TYPE(C_PTR) :: snow
    TYPE(C_PTR) :: snowh
    TYPE(C_PTR) :: canwat
    TYPE(C_PTR) :: isltyp
    TYPE(C_PTR) :: ivgtyp
    INTEGER(C_INT) :: isurban
    TYPE(C_PTR) :: tslb
    TYPE(C_PTR) :: smois
    TYPE(C_PTR) :: sh2o
    TYPE(C_PTR) :: dzs
    INTEGER(C_INT) :: fndsoilw
    INTEGER(C_INT) :: fndsnowh
    INTEGER(C_INT) :: isice
    INTEGER(C_INT) :: iswater
    TYPE(C_PTR) :: tsk
    TYPE(C_PTR) :: isnowxy
    TYPE(C_PTR) :: tvxy
    TYPE(C_PTR) :: tgxy
    TYPE(C_PTR) :: canicexy
    TYPE(C_PTR) :: tmn
    TYPE(C_PTR) :: xice
    TYPE(C_PTR) :: canliqxy
    TYPE(C_PTR) :: eahxy
    TYPE(C_PTR) :: tahxy
    TYPE(C_PTR) :: cmxy
    TYPE(C_PTR) :: chxy
    TYPE(C_PTR) :: fwetxy
    TYPE(C_PTR) :: sneqvoxy
    TYPE(C_PTR) :: alboldxy
    TYPE(C_PTR) :: qsnowxy
    TYPE(C_PTR) :: wslakexy
    TYPE(C_PTR) :: zwtxy
    TYPE(C_PTR) :: waxy
    TYPE(C_PTR) :: wtxy
    TYPE(C_PTR) :: tsnoxy
    TYPE(C_PTR) :: zsnsoxy
    TYPE(C_PTR) :: snicexy
    TYPE(C_PTR) :: snliqxy
    TYPE(C_PTR) :: lfmassxy
    TYPE(C_PTR) :: rtmassxy
    TYPE(C_PTR) :: stmassxy
    TYPE(C_PTR) :: woodxy
    TYPE(C_PTR) :: stblcpxy
    TYPE(C_PTR) :: fastcpxy
    TYPE(C_PTR) :: xsaixy
    TYPE(C_PTR) :: t2mvxy
    TYPE(C_PTR) :: t2mbxy
    TYPE(C_PTR) :: chstarxy
    INTEGER(C_INT) :: nsoil
    INTEGER(C_INT) :: restart
    INTEGER(C_INT) :: allowed_to_read
    INTEGER(C_INT) :: iopt_run
    INTEGER(C_INT) :: ids
    INTEGER(C_INT) :: ide
    INTEGER(C_INT) :: jds
    INTEGER(C_INT) :: jde
    INTEGER(C_INT) :: kds
    INTEGER(C_INT) :: kde
    INTEGER(C_INT) :: ims
    INTEGER(C_INT) :: ime
    INTEGER(C_INT) :: jms
    INTEGER(C_INT) :: jme
    INTEGER(C_INT) :: kms
    INTEGER(C_INT) :: kme
    INTEGER(C_INT) :: its
    INTEGER(C_INT) :: ite
    INTEGER(C_INT) :: jts
    INTEGER(C_INT) :: jte
    INTEGER(C_INT) :: kts
    INTEGER(C_INT) :: kte
    TYPE(C_PTR) :: smoiseq
    TYPE(C_PTR) :: smcwtdxy
    TYPE(C_PTR) :: rechxy
    TYPE(C_PTR) :: deeprechxy
    TYPE(C_PTR) :: areaxy
    REAL(C_FLOAT) :: dx
    REAL(C_FLOAT) :: dy
    TYPE(C_PTR) :: msftx
    TYPE(C_PTR) :: msfty
    REAL(C_FLOAT) :: wtddt
    TYPE(C_PTR) :: stepwtd
    REAL(C_FLOAT) :: dt
    TYPE(C_PTR) :: qrfsxy
    TYPE(C_PTR) :: qspringsxy
    TYPE(C_PTR) :: qslatxy
    TYPE(C_PTR) :: fdepthxy
    TYPE(C_PTR) :: ht
    TYPE(C_PTR) :: riverbedxy
    TYPE(C_PTR) :: eqzwt
    TYPE(C_PTR) :: rivercondxy
    TYPE(C_PTR) :: pexpxy
  END TYPE noahmp_init_args

  INTEGER(C_INT), PARAMETER :: NOAHMP_SYNC_FULL = 0, NOAHMP_SYNC_RESIDENT = 1
  INTEGER(C_INT), PARAMETER :: NOAHMP_HINT_DZ8W_CONSTANT = 1, NOAHMP_HINT_VEGFRA_UNCHANGED = 2, &
                               NOAHMP_HINT_P8W_LEVELS_EQUAL = 4
  INTEGER, PARAMETER :: NOAHMP_TABLES_BYTES = 12712   ! sizeof(noahmp_tables); checked at start-up

  TYPE(C_PTR), SAVE :: ctx = C_NULL_PTR
  INTEGER(C_INT8_T), SAVE, TARGET :: tables(NOAHMP_TABLES_BYTES)
  TYPE(noahmp_lsm_args), SAVE :: last_args
  INTEGER, SAVE :: global_nx = 0, global_ny = 0      ! extents of the whole domain when the run is tiled over GPUs
  INTEGER(C_INT), SAVE :: hints = 0
  LOGICAL, SAVE :: vegfra_new = .TRUE.

  INTERFACE
    FUNCTION noahmp_b200_read_tables(dir, dataset, soil, tbl) BIND(C, NAME="noahmp_b200_read_tables") RESULT(rc)
      IMPORT; CHARACTER(KIND=C_CHAR), INTENT(IN) :: dir(*), dataset(*), soil(*); TYPE(C_PTR), VALUE :: tbl
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION noahmp_b200_sizeof_tables() BIND(C, NAME="noahmp_b200_sizeof_tables") RESULT(n)
      IMPORT; INTEGER(C_LONG_LONG) :: n
    END FUNCTION
    FUNCTION noahmp_b200_create(device, tbl, ni, nj) BIND(C, NAME="noahmp_b200_create") RESULT(c)
      IMPORT; INTEGER(C_INT), VALUE :: device, ni, nj; TYPE(C_PTR), VALUE :: tbl; TYPE(C_PTR) :: c
    END FUNCTION
    SUBROUTINE noahmp_b200_destroy(c) BIND(C, NAME="noahmp_b200_destroy")
      IMPORT; TYPE(C_PTR), VALUE :: c
    END SUBROUTINE
    FUNCTION noahmp_b200_set_mode(c, mode) BIND(C, NAME="noahmp_b200_set_mode") RESULT(rc)
      IMPORT; TYPE(C_PTR), VALUE :: c; INTEGER(C_INT), VALUE :: mode; INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION noahmp_b200_set_fetch(c, fields) BIND(C, NAME="noahmp_b200_set_fetch") RESULT(rc)
      IMPORT; TYPE(C_PTR), VALUE :: c; CHARACTER(KIND=C_CHAR), INTENT(IN) :: fields(*); INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION noahmp_b200_noahmplsm(c, args, st) BIND(C, NAME="noahmp_b200_noahmplsm") RESULT(rc)
      IMPORT; TYPE(C_PTR), VALUE :: c; TYPE(noahmp_lsm_args), INTENT(IN) :: args; TYPE(noahmp_status) :: st
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION noahmp_b200_sync_host(c, args) BIND(C, NAME="noahmp_b200_sync_host") RESULT(rc)
      IMPORT; TYPE(C_PTR), VALUE :: c; TYPE(noahmp_lsm_args), INTENT(IN) :: args; INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION noahmp_b200_init(c, args) BIND(C, NAME="noahmp_b200_init") RESULT(rc)
      IMPORT; TYPE(C_PTR), VALUE :: c; TYPE(noahmp_init_args), INTENT(IN) :: args; INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION noahmp_b200_output_begin(c, args, fields, mask_water) BIND(C, NAME="noahmp_b200_output_begin") RESULT(rc)
      IMPORT; TYPE(C_PTR), VALUE :: c; TYPE(noahmp_lsm_args), INTENT(IN) :: args
      CHARACTER(KIND=C_CHAR), INTENT(IN) :: fields(*); INTEGER(C_INT), VALUE :: mask_water; INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION noahmp_b200_output_wait(c) BIND(C, NAME="noahmp_b200_output_wait") RESULT(rc)
      IMPORT; TYPE(C_PTR), VALUE :: c; INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION noahmp_b200_set_push(c, fields) BIND(C, NAME="noahmp_b200_set_push") RESULT(rc)
      IMPORT; TYPE(C_PTR), VALUE :: c; CHARACTER(KIND=C_CHAR), INTENT(IN) :: fields(*); INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION noahmp_b200_set_forcing_hints(c, h) BIND(C, NAME="noahmp_b200_set_forcing_hints") RESULT(rc)
      IMPORT; TYPE(C_PTR), VALUE :: c; INTEGER(C_INT), VALUE :: h; INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION noahmp_b200_get_status(c, st) BIND(C, NAME="noahmp_b200_get_status") RESULT(rc)
      IMPORT; TYPE(C_PTR), VALUE :: c; TYPE(noahmp_status) :: st; INTEGER(C_INT) :: rc
    END FUNCTION
    ! ---- opt_run = 5 groundwater and the communicator of the tiles
    FUNCTION noahmp_b200_wtable(c, args) BIND(C, NAME="noahmp_b200_wtable") RESULT(rc)
      IMPORT; TYPE(C_PTR), VALUE :: c; TYPE(noahmp_wtable_args), INTENT(IN) :: args; INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION noahmp_b200_wtable_begin(c, args) BIND(C, NAME="noahmp_b200_wtable_begin") RESULT(rc)
      IMPORT; TYPE(C_PTR), VALUE :: c; TYPE(noahmp_wtable_args), INTENT(IN) :: args; INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION noahmp_b200_wtable_exchange(c, stream) BIND(C, NAME="noahmp_b200_wtable_exchange") RESULT(rc)
      IMPORT; TYPE(C_PTR), VALUE :: c, stream; INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION noahmp_b200_wtable_end(c, args) BIND(C, NAME="noahmp_b200_wtable_end") RESULT(rc)
      IMPORT; TYPE(C_PTR), VALUE :: c; TYPE(noahmp_wtable_args), INTENT(IN) :: args; INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION noahmp_b200_wtable_sync_host(c, args) BIND(C, NAME="noahmp_b200_wtable_sync_host") RESULT(rc)
      IMPORT; TYPE(C_PTR), VALUE :: c; TYPE(noahmp_wtable_args), INTENT(IN) :: args; INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION noahmp_b200_comm_unique_id(id) BIND(C, NAME="noahmp_b200_comm_unique_id") RESULT(rc)
      IMPORT; TYPE(C_PTR), VALUE :: id; INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION noahmp_b200_comm_init(c, id, rank, nranks) BIND(C, NAME="noahmp_b200_comm_init") RESULT(rc)
      IMPORT; TYPE(C_PTR), VALUE :: c, id; INTEGER(C_INT), VALUE :: rank, nranks; INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION noahmp_b200_budget_enable(c, enable) BIND(C, NAME="noahmp_b200_budget_enable") RESULT(rc)
      IMPORT; TYPE(C_PTR), VALUE :: c; INTEGER(C_INT), VALUE :: enable; INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION noahmp_b200_budget_read(c, out8, global, reset) BIND(C, NAME="noahmp_b200_budget_read") RESULT(rc)
      IMPORT; TYPE(C_PTR), VALUE :: c; REAL(C_DOUBLE) :: out8(8); INTEGER(C_INT), VALUE :: global, reset
      INTEGER(C_INT) :: rc
    END FUNCTION
    ! ---- forcing pipeline on the device (row f2)
    FUNCTION c_forcing_static(c, lat2d, lon2d, zlvl) BIND(C, NAME="noahmp_b200_forcing_static") RESULT(rc)
      IMPORT; TYPE(C_PTR), VALUE :: c, lat2d, lon2d; REAL(C_FLOAT), VALUE :: zlvl; INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION c_forcing_upload(c, slot, f) BIND(C, NAME="noahmp_b200_forcing_upload") RESULT(rc)
      IMPORT; TYPE(C_PTR), VALUE :: c; INTEGER(C_INT), VALUE :: slot; TYPE(noahmp_forcing_fields), INTENT(IN) :: f
      INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION c_forcing_swap(c) BIND(C, NAME="noahmp_b200_forcing_swap") RESULT(rc)
      IMPORT; TYPE(C_PTR), VALUE :: c; INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION c_forcing_apply(c, fraction, iday, ihour, iminute, isecond, dt, julian) &
        BIND(C, NAME="noahmp_b200_forcing_apply") RESULT(rc)
      IMPORT; TYPE(C_PTR), VALUE :: c; REAL(C_FLOAT), VALUE :: fraction, dt
      INTEGER(C_INT), VALUE :: iday, ihour, iminute, isecond; REAL(C_FLOAT) :: julian; INTEGER(C_INT) :: rc
    END FUNCTION
    FUNCTION c_noahmplsm_device_forcing(c, args, st) BIND(C, NAME="noahmp_b200_noahmplsm_device_forcing") RESULT(rc)
      IMPORT; TYPE(C_PTR), VALUE :: c; TYPE(noahmp_lsm_args), INTENT(IN) :: args; TYPE(noahmp_status) :: st
      INTEGER(C_INT) :: rc
    END FUNCTION
  END INTERFACE

CONTAINS

  !> once, after read_mp_veg_parameters / SOIL_VEG_GEN_PARM would have run (they read the same files from the CWD)
  SUBROUTINE noahmp_b200_start(mminlu, ni, nj, device, resident)
    CHARACTER(LEN=*), INTENT(IN) :: mminlu
    INTEGER, INTENT(IN) :: ni, nj, device
    LOGICAL, INTENT(IN) :: resident
    INTEGER(C_INT) :: rc
    IF (noahmp_b200_sizeof_tables() /= NOAHMP_TABLES_BYTES) CALL wrf_error_fatal("noahmp_b200: table ABI mismatch")
    rc = noahmp_b200_read_tables("."//C_NULL_CHAR, TRIM(mminlu)//C_NULL_CHAR, "STAS"//C_NULL_CHAR, C_LOC(tables))
    IF (rc /= 0) CALL wrf_error_fatal("noahmp_b200: cannot read MPTABLE/VEGPARM/SOILPARM/GENPARM.TBL")
    ctx = noahmp_b200_create(INT(device, C_INT), C_LOC(tables), INT(ni, C_INT), INT(nj, C_INT))
    IF (.NOT. C_ASSOCIATED(ctx)) CALL wrf_error_fatal("noahmp_b200: no usable CUDA device (there is no CPU fallback)")
    IF (resident) THEN
      rc = noahmp_b200_set_mode(ctx, NOAHMP_SYNC_RESIDENT)
      rc = noahmp_b200_set_fetch(ctx, "tslb,xlaixy"//C_NULL_CHAR)   ! what land_driver_exe prints every step (:567-572)
      ! READFORC_HRLDAS overwrites LAI (= XLAIXY) from the forcing file before every call (:335, :403)
      rc = noahmp_b200_set_push(ctx, "xlaixy"//C_NULL_CHAR)
      ! DZ8W = 2*zlvl never changes (:344); P8W(:,2,:) = P8W(:,1,:) (:338); VEGFRA changes only when a forcing file
      ! brought a new one: the driver says so with noahmp_b200_vegfra_changed()
      hints = IOR(NOAHMP_HINT_DZ8W_CONSTANT, NOAHMP_HINT_P8W_LEVELS_EQUAL)
      rc = noahmp_b200_set_forcing_hints(ctx, hints)
    END IF
  END SUBROUTINE noahmp_b200_start

  !> Tiled run, one MPI rank per GPU: after noahmp_b200_start.  `id128` is the 128-byte NCCL id rank 0 obtained from
  !> noahmp_b200_comm_unique_id and the driver broadcast (MPI_Bcast(id128, 128, MPI_BYTE, 0, comm)); nx_global / ny_global
  !> are the extents of the whole domain.  From here on WTABLE_mmf_noahmp exchanges its halo over NCCL and
  !> noahmp_b200_global_budget sums over all tiles.
  SUBROUTINE noahmp_b200_start_parallel(id128, rank, nranks, nx_global, ny_global)
    INTEGER(C_INT8_T), INTENT(IN), TARGET :: id128(128)
    INTEGER, INTENT(IN) :: rank, nranks, nx_global, ny_global
    INTEGER(C_INT) :: rc
    rc = noahmp_b200_comm_init(ctx, C_LOC(id128), INT(rank, C_INT), INT(nranks, C_INT))
    IF (rc /= 0) CALL wrf_error_fatal("noahmp_b200: NCCL communicator could not be created")
    global_nx = nx_global
    global_ny = ny_global
    rc = noahmp_b200_budget_enable(ctx, 1_C_INT)
  END SUBROUTINE noahmp_b200_start_parallel

  !> call after READFORC_HRLDAS delivered a VEGFRA that differs from the previous one
  SUBROUTINE noahmp_b200_vegfra_changed()
    vegfra_new = .TRUE.
  END SUBROUTINE noahmp_b200_vegfra_changed

  !> eight fp64 sums over all tiles (one ncclAllReduce): storage, precipitation, ET, runoff (mm), energy residual
  !> (W/m2), SWE (mm), columns, steps — see include/noahmp_b200.h; reset starts a new reporting interval
  SUBROUTINE noahmp_b200_global_budget(sums, reset)
    REAL(C_DOUBLE), INTENT(OUT) :: sums(8)
    LOGICAL, INTENT(IN) :: reset
    INTEGER(C_INT) :: rc
    rc = noahmp_b200_budget_read(ctx, sums, 1_C_INT, MERGE(1_C_INT, 0_C_INT, reset))
    IF (rc /= 0) CALL wrf_error_fatal("noahmp_b200: budget_read failed (call noahmp_b200_start_parallel first)")
  END SUBROUTINE noahmp_b200_global_budget

  !> RESIDENT mode: call before hrldas_output / restart writes (driver :440-592) to refresh every host array
  SUBROUTINE noahmp_b200_refresh_host()
    INTEGER(C_INT) :: rc
    rc = noahmp_b200_sync_host(ctx, last_args)
    IF (rc /= 0) CALL wrf_error_fatal("noahmp_b200: sync_host failed")
  END SUBROUTINE noahmp_b200_refresh_host

  !> RESIDENT mode, asynchronous variant of refresh_host: snapshot the state as of the step just taken and copy it
  !> down while the next steps run.  history = .TRUE. masks water points with -1.E33 (what put_var_2d/3d would do).
  !> Call noahmp_b200_snapshot_wait() right before the nf90_put_var calls.
  SUBROUTINE noahmp_b200_snapshot_begin(history)
    LOGICAL, INTENT(IN) :: history
    INTEGER(C_INT) :: rc
    rc = noahmp_b200_output_begin(ctx, last_args, "*"//C_NULL_CHAR, MERGE(1_C_INT, 0_C_INT, history))
    IF (rc /= 0) CALL wrf_error_fatal("noahmp_b200: output_begin failed")
  END SUBROUTINE noahmp_b200_snapshot_begin

  SUBROUTINE noahmp_b200_snapshot_wait()
    INTEGER(C_INT) :: rc
    rc = noahmp_b200_output_wait(ctx)
    IF (rc /= 0) CALL wrf_error_fatal("noahmp_b200: output_wait failed")
  END SUBROUTINE noahmp_b200_snapshot_wait

  SUBROUTINE noahmp_b200_stop()
    CALL noahmp_b200_destroy(ctx)
    ctx = C_NULL_PTR
  END SUBROUTINE noahmp_b200_stop

  SUBROUTINE noahmplsm( &
      ITIMESTEP, YR, JULIAN, COSZIN, XLATIN, DZ8W, DT, DZS, NSOIL, DX, IVGTYP, ISLTYP, VEGFRA, &
      VEGMAX, TMN, XLAND, XICE, XICE_THRES, ISICE, ISURBAN, IDVEG, IOPT_CRS, IOPT_BTR, IOPT_RUN, &
      IOPT_SFC, IOPT_FRZ, IOPT_INF, IOPT_RAD, IOPT_ALB, IOPT_SNF, IOPT_TBOT, IOPT_STC, IZ0TLND, &
      T3D, QV3D, U_PHY, V_PHY, SWDOWN, GLW, P8W3D, RAINBL, TSK, HFX, QFX, LH, GRDFLX, SMSTAV, &
      SMSTOT, SFCRUNOFF, UDRUNOFF, ALBEDO, SNOWC, SMOIS, SH2O, TSLB, SNOW, SNOWH, CANWAT, ACSNOM, &
      ACSNOW, EMISS, QSFC, ISNOWXY, TVXY, TGXY, CANICEXY, CANLIQXY, EAHXY, TAHXY, CMXY, CHXY, &
      FWETXY, SNEQVOXY, ALBOLDXY, QSNOWXY, WSLAKEXY, ZWTXY, WAXY, WTXY, TSNOXY, ZSNSOXY, SNICEXY, &
      SNLIQXY, LFMASSXY, RTMASSXY, STMASSXY, WOODXY, STBLCPXY, FASTCPXY, XLAIXY, XSAIXY, TAUSSXY, &
      SMOISEQ, SMCWTDXY, DEEPRECHXY, RECHXY, T2MVXY, T2MBXY, Q2MVXY, Q2MBXY, TRADXY, NEEXY, GPPXY, &
      NPPXY, FVEGXY, RUNSFXY, RUNSBXY, ECANXY, EDIRXY, ETRANXY, FSAXY, FIRAXY, APARXY, PSNXY, &
      SAVXY, SAGXY, RSSUNXY, RSSHAXY, BGAPXY, WGAPXY, TGVXY, TGBXY, CHVXY, CHBXY, SHGXY, SHCXY, &
      SHBXY, EVGXY, EVBXY, GHVXY, GHBXY, IRGXY, IRCXY, IRBXY, TRXY, EVCXY, CHLEAFXY, CHUCXY, &
      CHV2XY, CHB2XY, IDS, IDE, JDS, JDE, KDS, KDE, IMS, IME, JMS, JME, KMS, KME, ITS, ITE, JTS, &
      JTE, KTS, KTE)
    IMPLICIT NONE
    INTEGER, INTENT(IN) :: ITIMESTEP
    INTEGER, INTENT(IN) :: YR
    REAL, INTENT(IN) :: JULIAN
    REAL, INTENT(IN) :: DT
    INTEGER, INTENT(IN) :: NSOIL
    REAL, INTENT(IN) :: DX
    REAL, INTENT(IN) :: XICE_THRES
    INTEGER, INTENT(IN) :: ISICE
    INTEGER, INTENT(IN) :: ISURBAN
    INTEGER, INTENT(IN) :: IDVEG
    INTEGER, INTENT(IN) :: IOPT_CRS
    INTEGER, INTENT(IN) :: IOPT_BTR
    INTEGER, INTENT(IN) :: IOPT_RUN
    INTEGER, INTENT(IN) :: IOPT_SFC
    INTEGER, INTENT(IN) :: IOPT_FRZ
    INTEGER, INTENT(IN) :: IOPT_INF
    INTEGER, INTENT(IN) :: IOPT_RAD
    INTEGER, INTENT(IN) :: IOPT_ALB
    INTEGER, INTENT(IN) :: IOPT_SNF
    INTEGER, INTENT(IN) :: IOPT_TBOT
    INTEGER, INTENT(IN) :: IOPT_STC
    INTEGER, INTENT(IN) :: IZ0TLND
    INTEGER, INTENT(IN) :: IDS
    INTEGER, INTENT(IN) :: IDE
    INTEGER, INTENT(IN) :: JDS
    INTEGER, INTENT(IN) :: JDE
    INTEGER, INTENT(IN) :: KDS
    INTEGER, INTENT(IN) :: KDE
    INTEGER, INTENT(IN) :: IMS
    INTEGER, INTENT(IN) :: IME
    INTEGER, INTENT(IN) :: JMS
    INTEGER, INTENT(IN) :: JME
    INTEGER, INTENT(IN) :: KMS
    INTEGER, INTENT(IN) :: KME
    INTEGER, INTENT(IN) :: ITS
    INTEGER, INTENT(IN) :: ITE
    INTEGER, INTENT(IN) :: JTS
    INTEGER, INTENT(IN) :: JTE
    INTEGER, INTENT(IN) :: KTS
    INTEGER, INTENT(IN) :: KTE
    REAL, INTENT(IN), TARGET :: COSZIN(ims:ime, jms:jme)
    REAL, INTENT(IN), TARGET :: XLATIN(ims:ime, jms:jme)
    REAL, INTENT(IN), TARGET :: DZ8W(ims:ime, kms:kme, jms:jme)
    REAL, INTENT(IN), TARGET :: DZS(1:NSOIL)
    INTEGER, INTENT(IN), TARGET :: IVGTYP(ims:ime, jms:jme)
    INTEGER, INTENT(IN), TARGET :: ISLTYP(ims:ime, jms:jme)
    REAL, INTENT(IN), TARGET :: VEGFRA(ims:ime, jms:jme)
    REAL, INTENT(IN), TARGET :: VEGMAX(ims:ime, jms:jme)
    REAL, INTENT(IN), TARGET :: TMN(ims:ime, jms:jme)
    REAL, INTENT(IN), TARGET :: XLAND(ims:ime, jms:jme)
    REAL, INTENT(IN), TARGET :: XICE(ims:ime, jms:jme)
    REAL, INTENT(IN), TARGET :: T3D(ims:ime, kms:kme, jms:jme)
    REAL, INTENT(IN), TARGET :: QV3D(ims:ime, kms:kme, jms:jme)
    REAL, INTENT(IN), TARGET :: U_PHY(ims:ime, kms:kme, jms:jme)
    REAL, INTENT(IN), TARGET :: V_PHY(ims:ime, kms:kme, jms:jme)
    REAL, INTENT(IN), TARGET :: SWDOWN(ims:ime, jms:jme)
    REAL, INTENT(IN), TARGET :: GLW(ims:ime, jms:jme)
    REAL, INTENT(IN), TARGET :: P8W3D(ims:ime, kms:kme, jms:jme)
    REAL, INTENT(IN), TARGET :: RAINBL(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: TSK(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: HFX(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: QFX(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: LH(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: GRDFLX(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: SMSTAV(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: SMSTOT(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: SFCRUNOFF(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: UDRUNOFF(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: ALBEDO(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: SNOWC(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: SMOIS(ims:ime, 1:NSOIL, jms:jme)
    REAL, INTENT(INOUT), TARGET :: SH2O(ims:ime, 1:NSOIL, jms:jme)
    REAL, INTENT(INOUT), TARGET :: TSLB(ims:ime, 1:NSOIL, jms:jme)
    REAL, INTENT(INOUT), TARGET :: SNOW(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: SNOWH(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: CANWAT(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: ACSNOM(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: ACSNOW(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: EMISS(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: QSFC(ims:ime, jms:jme)
    INTEGER, INTENT(INOUT), TARGET :: ISNOWXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: TVXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: TGXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: CANICEXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: CANLIQXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: EAHXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: TAHXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: CMXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: CHXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: FWETXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: SNEQVOXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: ALBOLDXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: QSNOWXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: WSLAKEXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: ZWTXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: WAXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: WTXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: TSNOXY(ims:ime, -2:0, jms:jme)
    REAL, INTENT(INOUT), TARGET :: ZSNSOXY(ims:ime, -2:NSOIL, jms:jme)
    REAL, INTENT(INOUT), TARGET :: SNICEXY(ims:ime, -2:0, jms:jme)
    REAL, INTENT(INOUT), TARGET :: SNLIQXY(ims:ime, -2:0, jms:jme)
    REAL, INTENT(INOUT), TARGET :: LFMASSXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: RTMASSXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: STMASSXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: WOODXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: STBLCPXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: FASTCPXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: XLAIXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: XSAIXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: TAUSSXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: SMOISEQ(ims:ime, 1:NSOIL, jms:jme)
    REAL, INTENT(INOUT), TARGET :: SMCWTDXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: DEEPRECHXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: RECHXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: T2MVXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: T2MBXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: Q2MVXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: Q2MBXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: TRADXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: NEEXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: GPPXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: NPPXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: FVEGXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: RUNSFXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: RUNSBXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: ECANXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: EDIRXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: ETRANXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: FSAXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: FIRAXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: APARXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: PSNXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: SAVXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: SAGXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: RSSUNXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: RSSHAXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: BGAPXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: WGAPXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: TGVXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: TGBXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: CHVXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: CHBXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: SHGXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: SHCXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: SHBXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: EVGXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: EVBXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: GHVXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: GHBXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: IRGXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: IRCXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: IRBXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: TRXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: EVCXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: CHLEAFXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: CHUCXY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: CHV2XY(ims:ime, jms:jme)
    REAL, INTENT(INOUT), TARGET :: CHB2XY(ims:ime, jms:jme)
    TYPE(noahmp_lsm_args) :: a
    TYPE(noahmp_status) :: st
    INTEGER(C_INT) :: rc
    CHARACTER(LEN=160) :: msg

    a%itimestep = ITIMESTEP
    a%yr = YR
    a%julian = JULIAN
    a%coszin = C_LOC(COSZIN)
    a%xlatin = C_LOC(XLATIN)
    a%dz8w = C_LOC(DZ8W)
    a%dt = DT
    a%dzs = C_LOC(DZS)
    a%nsoil = NSOIL
    a%dx = DX
    a%ivgtyp = C_LOC(IVGTYP)
    a%isltyp = C_LOC(ISLTYP)
    a%vegfra = C_LOC(VEGFRA)
    a%vegmax = C_LOC(VEGMAX)
    a%tmn = C_LOC(TMN)
    a%xland = C_LOC(XLAND)
    a%xice = C_LOC(XICE)
    a%xice_thres = XICE_THRES
    a%isice = ISICE
    a%isurban = ISURBAN
    a%idveg = IDVEG
    a%iopt_crs = IOPT_CRS
    a%iopt_btr = IOPT_BTR
    a%iopt_run = IOPT_RUN
    a%iopt_sfc = IOPT_SFC
    a%iopt_frz = IOPT_FRZ
    a%iopt_inf = IOPT_INF
    a%iopt_rad = IOPT_RAD
    a%iopt_alb = IOPT_ALB
    a%iopt_snf = IOPT_SNF
    a%iopt_tbot = IOPT_TBOT
    a%iopt_stc = IOPT_STC
    a%iz0tlnd = IZ0TLND
    a%t3d = C_LOC(T3D)
    a%qv3d = C_LOC(QV3D)
    a%u_phy = C_LOC(U_PHY)
    a%v_phy = C_LOC(V_PHY)
    a%swdown = C_LOC(SWDOWN)
    a%glw = C_LOC(GLW)
    a%p8w3d = C_LOC(P8W3D)
    a%rainbl = C_LOC(RAINBL)
    a%tsk = C_LOC(TSK)
    a%hfx = C_LOC(HFX)
    a%qfx = C_LOC(QFX)
    a%lh = C_LOC(LH)
    a%grdflx = C_LOC(GRDFLX)
    a%smstav = C_LOC(SMSTAV)
    a%smstot = C_LOC(SMSTOT)
    a%sfcrunoff = C_LOC(SFCRUNOFF)
    a%udrunoff = C_LOC(UDRUNOFF)
    a%albedo = C_LOC(ALBEDO)
    a%snowc = C_LOC(SNOWC)
    a%smois = C_LOC(SMOIS)
    a%sh2o = C_LOC(SH2O)
    a%tslb = C_LOC(TSLB)
    a%snow = C_LOC(SNOW)
    a%snowh = C_LOC(SNOWH)
    a%canwat = C_LOC(CANWAT)
    a%acsnom = C_LOC(ACSNOM)
    a%acsnow = C_LOC(ACSNOW)
    a%emiss = C_LOC(EMISS)
    a%qsfc = C_LOC(QSFC)
    a%isnowxy = C_LOC(ISNOWXY)
    a%tvxy = C_LOC(TVXY)
    a%tgxy = C_LOC(TGXY)
    a%canicexy = C_LOC(CANICEXY)
    a%canliqxy = C_LOC(CANLIQXY)
    a%eahxy = C_LOC(EAHXY)
    a%tahxy = C_LOC(TAHXY)
    a%cmxy = C_LOC(CMXY)
    a%chxy = C_LOC(CHXY)
    a%fwetxy = C_LOC(FWETXY)
    a%sneqvoxy = C_LOC(SNEQVOXY)
    a%alboldxy = C_LOC(ALBOLDXY)
    a%qsnowxy = C_LOC(QSNOWXY)
    a%wslakexy = C_LOC(WSLAKEXY)
    a%zwtxy = C_LOC(ZWTXY)
    a%waxy = C_LOC(WAXY)
    a%wtxy = C_LOC(WTXY)
    a%tsnoxy = C_LOC(TSNOXY)
    a%zsnsoxy = C_LOC(ZSNSOXY)
    a%snicexy = C_LOC(SNICEXY)
    a%snliqxy = C_LOC(SNLIQXY)
    a%lfmassxy = C_LOC(LFMASSXY)
    a%rtmassxy = C_LOC(RTMASSXY)
    a%stmassxy = C_LOC(STMASSXY)
    a%woodxy = C_LOC(WOODXY)
    a%stblcpxy = C_LOC(STBLCPXY)
    a%fastcpxy = C_LOC(FASTCPXY)
    a%xlaixy = C_LOC(XLAIXY)
    a%xsaixy = C_LOC(XSAIXY)
    a%taussxy = C_LOC(TAUSSXY)
    a%smoiseq = C_LOC(SMOISEQ)
    a%smcwtdxy = C_LOC(SMCWTDXY)
    a%deeprechxy = C_LOC(DEEPRECHXY)
    a%rechxy = C_LOC(RECHXY)
    a%t2mvxy = C_LOC(T2MVXY)
    a%t2mbxy = C_LOC(T2MBXY)
    a%q2mvxy = C_LOC(Q2MVXY)
    a%q2mbxy = C_LOC(Q2MBXY)
    a%tradxy = C_LOC(TRADXY)
    a%neexy = C_LOC(NEEXY)
    a%gppxy = C_LOC(GPPXY)
    a%nppxy = C_LOC(NPPXY)
    a%fvegxy = C_LOC(FVEGXY)
    a%runsfxy = C_LOC(RUNSFXY)
    a%runsbxy = C_LOC(RUNSBXY)
    a%ecanxy = C_LOC(ECANXY)
    a%edirxy = C_LOC(EDIRXY)
    a%etranxy = C_LOC(ETRANXY)
    a%fsaxy = C_LOC(FSAXY)
    a%firaxy = C_LOC(FIRAXY)
    a%aparxy = C_LOC(APARXY)
    a%psnxy = C_LOC(PSNXY)
    a%savxy = C_LOC(SAVXY)
    a%sagxy = C_LOC(SAGXY)
    a%rssunxy = C_LOC(RSSUNXY)
    a%rsshaxy = C_LOC(RSSHAXY)
    a%bgapxy = C_LOC(BGAPXY)
    a%wgapxy = C_LOC(WGAPXY)
    a%tgvxy = C_LOC(TGVXY)
    a%tgbxy = C_LOC(TGBXY)
    a%chvxy = C_LOC(CHVXY)
    a%chbxy = C_LOC(CHBXY)
    a%shgxy = C_LOC(SHGXY)
    a%shcxy = C_LOC(SHCXY)
    a%shbxy = C_LOC(SHBXY)
    a%evgxy = C_LOC(EVGXY)
    a%evbxy = C_LOC(EVBXY)
    a%ghvxy = C_LOC(GHVXY)
    a%ghbxy = C_LOC(GHBXY)
    a%irgxy = C_LOC(IRGXY)
    a%ircxy = C_LOC(IRCXY)
    a%irbxy = C_LOC(IRBXY)
    a%trxy = C_LOC(TRXY)
    a%evcxy = C_LOC(EVCXY)
    a%chleafxy = C_LOC(CHLEAFXY)
    a%chucxy = C_LOC(CHUCXY)
    a%chv2xy = C_LOC(CHV2XY)
    a%chb2xy = C_LOC(CHB2XY)
    a%ids = IDS
    a%ide = IDE
    a%jds = JDS
    a%jde = JDE
    a%kds = KDS
    a%kde = KDE
    a%ims = IMS
    a%ime = IME
    a%jms = JMS
    a%jme = JME
    a%kms = KMS
    a%kme = KME
    a%its = ITS
    a%ite = ITE
    a%jts = JTS
    a%jte = JTE
    a%kts = KTS
    a%kte = KTE
    last_args = a
    IF (hints /= 0) THEN
      rc = noahmp_b200_set_forcing_hints(ctx, MERGE(hints, IOR(hints, NOAHMP_HINT_VEGFRA_UNCHANGED), vegfra_new))
      vegfra_new = .FALSE.
    END IF
    rc = noahmp_b200_noahmplsm(ctx, a, st)
    IF (rc /= 0) THEN
      ! same fatal convention as the reference (util/module_wrf_utilities.F:12-24), same message texts
      SELECT CASE (st%code)
      CASE (1); WRITE(msg, '(A,2I6,ES14.6)') "Stop in Noah-MP (ERRSW) at i,j: ", st%i, st%j, st%value
      CASE (2); WRITE(msg, '(A,2I6,ES14.6)') "Energy budget problem in NOAHMP LSM at i,j: ", st%i, st%j, st%value
      CASE (3); WRITE(msg, '(A,2I6,ES14.6)') "Water budget problem in NOAHMP LSM at i,j: ", st%i, st%j, st%value
      CASE (4); WRITE(msg, '(A,2I6)') "STOP in Noah-MP: emitted longwave <0 at i,j: ", st%i, st%j
      CASE (5); WRITE(msg, '(A,2I6)') "CRITICAL PROBLEM: HCAN <= ZPD at i,j: ", st%i, st%j
      CASE (6); WRITE(msg, '(A,2I6)') "STOP in Noah-MP: ZLVL <= ZPD at i,j: ", st%i, st%j
      CASE (7); WRITE(msg, '(A,2I6)') "Warning: too many input soil/landuse types or NROOT at i,j: ", st%i, st%j
      CASE DEFAULT; WRITE(msg, '(A,I6)') "noahmp_b200 failure, code ", rc
      END SELECT
      CALL wrf_error_fatal(TRIM(msg))
    END IF
  END SUBROUTINE noahmplsm

  !> Drop-in for NOAHMP_INIT (phys/module_sf_noahmpdrv.F90:847-1179): same name, dummy list and order.  The tables are
  !> those noahmp_b200_start() read; call it after noahmp_b200_start().
  SUBROUTINE NOAHMP_INIT( &
      MMINLU, SNOW, SNOWH, CANWAT, ISLTYP, IVGTYP, ISURBAN, TSLB, SMOIS, SH2O, DZS, FNDSOILW, FNDSNOWH, &
      ISICE, ISWATER, TSK, ISNOWXY, TVXY, TGXY, CANICEXY, TMN, XICE, CANLIQXY, EAHXY, TAHXY, CMXY, CHXY, &
      FWETXY, SNEQVOXY, ALBOLDXY, QSNOWXY, WSLAKEXY, ZWTXY, WAXY, WTXY, TSNOXY, ZSNSOXY, SNICEXY, SNLIQXY, &
      LFMASSXY, RTMASSXY, STMASSXY, WOODXY, STBLCPXY, FASTCPXY, XSAIXY, T2MVXY, T2MBXY, CHSTARXY, NSOIL, &
      RESTART, ALLOWED_TO_READ, IOPT_RUN, IDS, IDE, JDS, JDE, KDS, KDE, IMS, IME, JMS, JME, KMS, KME, ITS, &
      ITE, JTS, JTE, KTS, KTE, SMOISEQ, SMCWTDXY, RECHXY, DEEPRECHXY, AREAXY, DX, DY, MSFTX, MSFTY, WTDDT, &
      STEPWTD, DT, QRFSXY, QSPRINGSXY, QSLATXY, FDEPTHXY, HT, RIVERBEDXY, EQZWT, RIVERCONDXY, PEXPXY)
    CHARACTER(LEN=*), INTENT(IN) :: MMINLU
    INTEGER, INTENT(IN) :: ids,ide, jds,jde, kds,kde, ims,ime, jms,jme, kms,kme, its,ite, jts,jte, kts,kte
    INTEGER, INTENT(IN) :: NSOIL, ISICE, ISWATER, ISURBAN, iopt_run
    LOGICAL, INTENT(IN) :: restart, allowed_to_read, FNDSOILW, FNDSNOWH
    REAL, DIMENSION(NSOIL), INTENT(IN), TARGET :: DZS
    REAL, INTENT(IN), OPTIONAL :: DX, DY, DT, WTDDT
    INTEGER, INTENT(OUT), OPTIONAL, TARGET :: STEPWTD
    INTEGER, DIMENSION(ims:ime,jms:jme), INTENT(IN), TARGET :: ISLTYP, IVGTYP
    INTEGER, DIMENSION(ims:ime,jms:jme), INTENT(INOUT), TARGET :: isnowxy
    REAL, DIMENSION(ims:ime,jms:jme), INTENT(IN), TARGET :: TSK, XICE
    REAL, DIMENSION(ims:ime,jms:jme), INTENT(INOUT), TARGET :: TMN
    REAL, DIMENSION(ims:ime,jms:jme), INTENT(INOUT), TARGET :: &
         snow, snowh, canwat, tvxy, tgxy, canicexy, canliqxy, eahxy, tahxy, cmxy, chxy, fwetxy, sneqvoxy, alboldxy, &
         qsnowxy, wslakexy, zwtxy, waxy, wtxy, lfmassxy, rtmassxy, stmassxy, woodxy, stblcpxy, fastcpxy, xsaixy, t2mvxy, t2mbxy, chstarxy
    REAL, DIMENSION(ims:ime,NSOIL,jms:jme), INTENT(INOUT), TARGET :: tslb
    REAL, DIMENSION(ims:ime,NSOIL,jms:jme), INTENT(INOUT), TARGET :: smois
    REAL, DIMENSION(ims:ime,NSOIL,jms:jme), INTENT(INOUT), TARGET :: sh2o
    REAL, DIMENSION(ims:ime,-2:0,jms:jme), INTENT(INOUT), TARGET :: tsnoxy
    REAL, DIMENSION(ims:ime,-2:0,jms:jme), INTENT(INOUT), TARGET :: snicexy
    REAL, DIMENSION(ims:ime,-2:0,jms:jme), INTENT(INOUT), TARGET :: snliqxy
    REAL, DIMENSION(ims:ime,-2:NSOIL,jms:jme), INTENT(INOUT), TARGET :: zsnsoxy
    REAL, DIMENSION(ims:ime,1:NSOIL,jms:jme), INTENT(INOUT), OPTIONAL, TARGET :: smoiseq
    REAL, DIMENSION(ims:ime,jms:jme), INTENT(INOUT), OPTIONAL, TARGET :: smcwtdxy, rechxy, deeprechxy, areaxy, qrfsxy, qspringsxy, qslatxy
    REAL, DIMENSION(ims:ime,jms:jme), INTENT(IN), OPTIONAL, TARGET :: msftx, msfty, fdepthxy, ht, riverbedxy, eqzwt, rivercondxy, pexpxy
    TYPE(noahmp_init_args) :: a
    INTEGER(C_INT) :: rc

    a%snow = C_LOC(snow)
    a%snowh = C_LOC(snowh)
    a%canwat = C_LOC(canwat)
    a%isltyp = C_LOC(isltyp)
    a%ivgtyp = C_LOC(ivgtyp)
    a%isurban = isurban
    a%tslb = C_LOC(tslb)
    a%smois = C_LOC(smois)
    a%sh2o = C_LOC(sh2o)
    a%dzs = C_LOC(dzs)
    a%fndsoilw = MERGE(1_C_INT, 0_C_INT, fndsoilw)
    a%fndsnowh = MERGE(1_C_INT, 0_C_INT, fndsnowh)
    a%isice = isice
    a%iswater = iswater
    a%tsk = C_LOC(tsk)
    a%isnowxy = C_LOC(isnowxy)
    a%tvxy = C_LOC(tvxy)
    a%tgxy = C_LOC(tgxy)
    a%canicexy = C_LOC(canicexy)
    a%tmn = C_LOC(tmn)
    a%xice = C_LOC(xice)
    a%canliqxy = C_LOC(canliqxy)
    a%eahxy = C_LOC(eahxy)
    a%tahxy = C_LOC(tahxy)
    a%cmxy = C_LOC(cmxy)
    a%chxy = C_LOC(chxy)
    a%fwetxy = C_LOC(fwetxy)
    a%sneqvoxy = C_LOC(sneqvoxy)
    a%alboldxy = C_LOC(alboldxy)
    a%qsnowxy = C_LOC(qsnowxy)
    a%wslakexy = C_LOC(wslakexy)
    a%zwtxy = C_LOC(zwtxy)
    a%waxy = C_LOC(waxy)
    a%wtxy = C_LOC(wtxy)
    a%tsnoxy = C_LOC(tsnoxy)
    a%zsnsoxy = C_LOC(zsnsoxy)
    a%snicexy = C_LOC(snicexy)
    a%snliqxy = C_LOC(snliqxy)
    a%lfmassxy = C_LOC(lfmassxy)
    a%rtmassxy = C_LOC(rtmassxy)
    a%stmassxy = C_LOC(stmassxy)
    a%woodxy = C_LOC(woodxy)
    a%stblcpxy = C_LOC(stblcpxy)
    a%fastcpxy = C_LOC(fastcpxy)
    a%xsaixy = C_LOC(xsaixy)
    a%t2mvxy = C_LOC(t2mvxy)
    a%t2mbxy = C_LOC(t2mbxy)
    a%chstarxy = C_LOC(chstarxy)
    a%nsoil = nsoil
    a%restart = MERGE(1_C_INT, 0_C_INT, restart)
    a%allowed_to_read = MERGE(1_C_INT, 0_C_INT, allowed_to_read)
    a%iopt_run = iopt_run
    a%ids = ids
    a%ide = ide
    a%jds = jds
    a%jde = jde
    a%kds = kds
    a%kde = kde
    a%ims = ims
    a%ime = ime
    a%jms = jms
    a%jme = jme
    a%kms = kms
    a%kme = kme
    a%its = its
    a%ite = ite
    a%jts = jts
    a%jte = jte
    a%kts = kts
    a%kte = kte
    ! optional groundwater block: absent arguments travel as NULL and the library answers as the reference does
    ! ('Not enough fields to use groundwater option in Noah-MP') when iopt_run = 5 needs them
    a%smoiseq = C_NULL_PTR; IF (PRESENT(smoiseq)) a%smoiseq = C_LOC(smoiseq)
    a%smcwtdxy = C_NULL_PTR; IF (PRESENT(smcwtdxy)) a%smcwtdxy = C_LOC(smcwtdxy)
    a%rechxy = C_NULL_PTR; IF (PRESENT(rechxy)) a%rechxy = C_LOC(rechxy)
    a%deeprechxy = C_NULL_PTR; IF (PRESENT(deeprechxy)) a%deeprechxy = C_LOC(deeprechxy)
    a%areaxy = C_NULL_PTR; IF (PRESENT(areaxy)) a%areaxy = C_LOC(areaxy)
    a%msftx = C_NULL_PTR; IF (PRESENT(msftx)) a%msftx = C_LOC(msftx)
    a%msfty = C_NULL_PTR; IF (PRESENT(msfty)) a%msfty = C_LOC(msfty)
    a%stepwtd = C_NULL_PTR; IF (PRESENT(stepwtd)) a%stepwtd = C_LOC(stepwtd)
    a%qrfsxy = C_NULL_PTR; IF (PRESENT(qrfsxy)) a%qrfsxy = C_LOC(qrfsxy)
    a%qspringsxy = C_NULL_PTR; IF (PRESENT(qspringsxy)) a%qspringsxy = C_LOC(qspringsxy)
    a%qslatxy = C_NULL_PTR; IF (PRESENT(qslatxy)) a%qslatxy = C_LOC(qslatxy)
    a%fdepthxy = C_NULL_PTR; IF (PRESENT(fdepthxy)) a%fdepthxy = C_LOC(fdepthxy)
    a%ht = C_NULL_PTR; IF (PRESENT(ht)) a%ht = C_LOC(ht)
    a%riverbedxy = C_NULL_PTR; IF (PRESENT(riverbedxy)) a%riverbedxy = C_LOC(riverbedxy)
    a%eqzwt = C_NULL_PTR; IF (PRESENT(eqzwt)) a%eqzwt = C_LOC(eqzwt)
    a%rivercondxy = C_NULL_PTR; IF (PRESENT(rivercondxy)) a%rivercondxy = C_LOC(rivercondxy)
    a%pexpxy = C_NULL_PTR; IF (PRESENT(pexpxy)) a%pexpxy = C_LOC(pexpxy)
    a%dx = 0.0; IF (PRESENT(dx)) a%dx = dx
    a%dy = 0.0; IF (PRESENT(dy)) a%dy = dy
    a%wtddt = 0.0; IF (PRESENT(wtddt)) a%wtddt = wtddt
    a%dt = 0.0; IF (PRESENT(dt)) a%dt = dt
    rc = noahmp_b200_init(ctx, a)
    IF (rc == 9) CALL wrf_error_fatal("module_sf_noahlsm.F: lsminit: out of range value of ISLTYP. Is this field in the input?")
    IF (rc /= 0 .AND. iopt_run == 5) CALL wrf_error_fatal('Not enough fields to use groundwater option in Noah-MP')
    IF (rc /= 0) CALL wrf_error_fatal("noahmp_b200_init failed")
  END SUBROUTINE NOAHMP_INIT

  !> Drop-in for WTABLE_mmf_noahmp (phys/module_sf_noahmp_groundwater.F90:14-22; call site
  !> driver/module_hrldas_noahmp_driver.F90:420-436): same name, dummy list and order.  In a tiled run the reference's
  !> MPI build passes ids = its ... and never exchanges a halo, so its answer depends on the rank count; here the domain
  !> extents given to noahmp_b200_start_parallel replace ids..jde and the library exchanges the KCELL / HEAD halo over
  !> NCCL inside the call: every tiling reproduces the sequential single-domain result.
  SUBROUTINE WTABLE_mmf_noahmp( &
      NSOIL, XLAND, XICE, XICE_THRESHOLD, ISICE, ISLTYP, SMOISEQ, DZS, WTDDT, FDEPTH, AREA, TOPO, ISURBAN, IVGTYP, &
      RIVERCOND, RIVERBED, EQWTD, PEXP, SMOIS, SH2OXY, SMCWTD, WTD, QRF, DEEPRECH, QSPRING, QSLAT, QRFS, QSPRINGS, &
      RECH, IDS, IDE, JDS, JDE, KDS, KDE, IMS, IME, JMS, JME, KMS, KME, ITS, ITE, JTS, JTE, KTS, KTE)
    IMPLICIT NONE
    INTEGER, INTENT(IN) :: IDS, IDE, JDS, JDE, KDS, KDE, IMS, IME, JMS, JME, KMS, KME, ITS, ITE, JTS, JTE, KTS, KTE
    REAL, INTENT(IN) :: WTDDT
    REAL, INTENT(IN) :: XICE_THRESHOLD
    INTEGER, INTENT(IN) :: ISICE
    REAL, DIMENSION(ims:ime, jms:jme), INTENT(IN), TARGET :: XLAND, XICE
    INTEGER, DIMENSION(ims:ime, jms:jme), INTENT(IN), TARGET :: ISLTYP, IVGTYP
    INTEGER, INTENT(IN) :: NSOIL
    INTEGER, INTENT(IN) :: ISURBAN
    REAL, DIMENSION(ims:ime, 1:nsoil, jms:jme), INTENT(IN), TARGET :: SMOISEQ
    REAL, DIMENSION(1:nsoil), INTENT(IN), TARGET :: DZS
    REAL, DIMENSION(ims:ime, jms:jme), INTENT(IN), TARGET :: FDEPTH, AREA, TOPO, EQWTD, PEXP, RIVERBED, RIVERCOND
    REAL, DIMENSION(ims:ime, 1:nsoil, jms:jme), INTENT(INOUT), TARGET :: SMOIS, SH2OXY
    REAL, DIMENSION(ims:ime, jms:jme), INTENT(INOUT), TARGET :: WTD, SMCWTD, DEEPRECH, QSLAT, QRFS, QSPRINGS, RECH
    REAL, DIMENSION(ims:ime, jms:jme), INTENT(OUT), TARGET :: QRF, QSPRING
    TYPE(noahmp_wtable_args) :: a
    INTEGER(C_INT) :: rc

    a%nsoil = NSOIL
    a%xland = C_LOC(XLAND)
    a%xice = C_LOC(XICE)
    a%xice_threshold = XICE_THRESHOLD
    a%isice = ISICE
    a%isltyp = C_LOC(ISLTYP)
    a%smoiseq = C_LOC(SMOISEQ)
    a%dzs = C_LOC(DZS)
    a%wtddt = WTDDT
    a%fdepth = C_LOC(FDEPTH)
    a%area = C_LOC(AREA)
    a%topo = C_LOC(TOPO)
    a%isurban = ISURBAN
    a%ivgtyp = C_LOC(IVGTYP)
    a%rivercond = C_LOC(RIVERCOND)
    a%riverbed = C_LOC(RIVERBED)
    a%eqwtd = C_LOC(EQWTD)
    a%pexp = C_LOC(PEXP)
    a%smois = C_LOC(SMOIS)
    a%sh2oxy = C_LOC(SH2OXY)
    a%smcwtd = C_LOC(SMCWTD)
    a%wtd = C_LOC(WTD)
    a%qrf = C_LOC(QRF)
    a%deeprech = C_LOC(DEEPRECH)
    a%qspring = C_LOC(QSPRING)
    a%qslat = C_LOC(QSLAT)
    a%qrfs = C_LOC(QRFS)
    a%qsprings = C_LOC(QSPRINGS)
    a%rech = C_LOC(RECH)
    a%ids = IDS
    a%ide = IDE
    a%jds = JDS
    a%jde = JDE
    a%kds = KDS
    a%kde = KDE
    a%ims = IMS
    a%ime = IME
    a%jms = JMS
    a%jme = JME
    a%kms = KMS
    a%kme = KME
    a%its = ITS
    a%ite = ITE
    a%jts = JTS
    a%jte = JTE
    a%kts = KTS
    a%kte = KTE
    IF (global_nx > 0) THEN   ! tiled run: the stencil and its clipping refer to the whole domain
      a%ids = 1
      a%ide = global_nx
      a%jds = 1
      a%jde = global_ny
    END IF
    rc = noahmp_b200_wtable(ctx, a)
    IF (rc /= 0) CALL wrf_error_fatal("noahmp_b200_wtable failed")
  END SUBROUTINE WTABLE_mmf_noahmp

  ! ---- forcing pipeline on the device (INTEGRATION.md section 4) ---------------------------------------------------
  SUBROUTINE noahmp_b200_forcing_static(lat2d, lon2d, zlvl)
    REAL, INTENT(IN), TARGET, CONTIGUOUS :: lat2d(:,:), lon2d(:,:)
    REAL, INTENT(IN) :: zlvl
    INTEGER(C_INT) :: rc
    rc = c_forcing_static(ctx, C_LOC(lat2d), C_LOC(lon2d), REAL(zlvl, C_FLOAT))
    IF (rc /= 0) CALL wrf_error_fatal("noahmp_b200_forcing_static failed")
  END SUBROUTINE noahmp_b200_forcing_static

  SUBROUTINE noahmp_b200_forcing_upload(slot, f)
    INTEGER, INTENT(IN) :: slot
    TYPE(noahmp_forcing_fields), INTENT(IN) :: f
    INTEGER(C_INT) :: rc
    rc = c_forcing_upload(ctx, INT(slot, C_INT), f)
    IF (rc /= 0) CALL wrf_error_fatal("noahmp_b200_forcing_upload failed")
  END SUBROUTINE noahmp_b200_forcing_upload

  SUBROUTINE noahmp_b200_forcing_swap()
    INTEGER(C_INT) :: rc
    rc = c_forcing_swap(ctx)
  END SUBROUTINE noahmp_b200_forcing_swap

  SUBROUTINE noahmp_b200_forcing_apply(fraction, iday, ihour, iminute, isecond, dt, julian)
    REAL, INTENT(IN) :: fraction, dt
    INTEGER, INTENT(IN) :: iday, ihour, iminute, isecond
    REAL, INTENT(OUT) :: julian
    REAL(C_FLOAT) :: j
    INTEGER(C_INT) :: rc
    rc = c_forcing_apply(ctx, REAL(fraction, C_FLOAT), INT(iday, C_INT), INT(ihour, C_INT), INT(iminute, C_INT), &
                         INT(isecond, C_INT), REAL(dt, C_FLOAT), j)
    IF (rc /= 0) CALL wrf_error_fatal("noahmp_b200_forcing_apply failed")
    julian = j
  END SUBROUTINE noahmp_b200_forcing_apply

  !> the step with the forcing planes noahmp_b200_forcing_apply filled: call noahmplsm once (it records its argument
  !> list), then this routine for the following steps
  SUBROUTINE noahmplsm_device_forcing(itimestep, yr, julian)
    INTEGER, INTENT(IN) :: itimestep, yr
    REAL, INTENT(IN) :: julian
    TYPE(noahmp_status) :: st
    INTEGER(C_INT) :: rc
    last_args%itimestep = itimestep
    last_args%yr = yr
    last_args%julian = julian
    rc = c_noahmplsm_device_forcing(ctx, last_args, st)
    IF (rc /= 0) CALL wrf_error_fatal("noahmp_b200: model check failed in noahmplsm_device_forcing")
  END SUBROUTINE noahmplsm_device_forcing

END MODULE module_sf_noahmpdrv_b200
