! Test program of oracle/ref/f90cxx.py (the Fortran -> C++ translator of the reference pin): one small routine per
! language rule the Noah-MP sources rely on.  tests/test_f90cxx.py translates this file, compiles it with g++ and
! checks every result against the value the Fortran standard (and gfortran, where the standard leaves a choice)
! prescribes.  Nothing here comes from the reference.
MODULE SEM_CONSTANTS
  IMPLICIT NONE
  REAL, PARAMETER :: CICE = 2106.        ! hides SEM_GLOBALS' CICE where this module is USEd inside a routine
  REAL, PARAMETER :: THIRD = 1./3.
END MODULE SEM_CONSTANTS

MODULE SEM_GLOBALS
  IMPLICIT NONE
  REAL, PARAMETER :: CICE = 2.094E06
  INTEGER, PARAMETER :: NB = 2
  REAL :: SHARED_SCALAR
  REAL, DIMENSION(1:NB) :: SHARED_VEC
  REAL :: TAB2(3,NB)
  DATA (TAB2(I,1),I=1,3) /1.0, 2.0, 3.0/
  DATA (TAB2(I,2),I=1,3) /10.0, 20.0, 30.0/
  INTEGER :: I
END MODULE SEM_GLOBALS

MODULE SEM_ROUTINES
  USE SEM_GLOBALS
  IMPLICIT NONE
CONTAINS

  SUBROUTINE PRECEDENCE (A, B, C, OUT)
    REAL, INTENT(IN) :: A, B, C
    REAL, DIMENSION(8), INTENT(OUT) :: OUT
    OUT(1) = -A**2                 ! -(A**2)
    OUT(2) = A - B + C             ! (A-B)+C
    OUT(3) = A / B * C             ! (A/B)*C
    OUT(4) = 2.**3**2              ! 2**(3**2) = 512
    OUT(5) = -A * B                ! -(A*B)
    OUT(6) = A + B * C ** 2
    OUT(7) = (A + B) * C
    OUT(8) = A ** (-2)
  END SUBROUTINE PRECEDENCE

  SUBROUTINE INTEGER_RULES (N, M, OUT)
    INTEGER, INTENT(IN) :: N, M
    REAL, DIMENSION(8), INTENT(OUT) :: OUT
    INTEGER :: K
    OUT(1) = N / M                 ! integer division, truncated toward zero
    OUT(2) = (-N) / M
    OUT(3) = 1 / 2 * 3.0           ! (1/2)*3.0 = 0
    OUT(4) = REAL(N) / M
    OUT(5) = MOD(-N, M)            ! sign of the first argument
    OUT(6) = NINT(2.5) + NINT(-2.5)   ! halves away from zero: 3 + (-3)
    OUT(7) = INT(-2.7)
    K = 2.9                        ! assignment truncates
    OUT(8) = K
  END SUBROUTINE INTEGER_RULES

  SUBROUTINE LOWER_BOUNDS (NSNOW, NSOIL, ISNOW, Z, OUT)
    INTEGER, INTENT(IN) :: NSNOW, NSOIL, ISNOW
    REAL, DIMENSION(-NSNOW+1:NSOIL), INTENT(INOUT) :: Z
    REAL, DIMENSION(6), INTENT(OUT) :: OUT
    REAL, DIMENSION(-NSNOW+1:0) :: W
    INTEGER :: K
    W = 7.
    W(ISNOW+1:0) = Z(ISNOW+1:0) * 2.
    OUT(1) = W(-NSNOW+1)
    OUT(2) = W(0)
    DO K = -NSNOW+1, NSOIL
       Z(K) = Z(K) + K
    END DO
    OUT(3) = Z(-NSNOW+1)
    OUT(4) = Z(NSOIL)
    OUT(5) = MAXVAL(Z)
    OUT(6) = 0.
    IF (ANY(Z(1:NSOIL) > 100.) .AND. ANY(Z(1:NSOIL) < 100.)) OUT(6) = 1.
  END SUBROUTINE LOWER_BOUNDS

  SUBROUTINE DO_LOOPS (N, OUT)
    INTEGER, INTENT(IN) :: N
    REAL, DIMENSION(8), INTENT(OUT) :: OUT
    INTEGER :: I, J, M, CNT
    M = N
    CNT = 0
    DO I = 1, M                    ! the trip count is fixed at entry although M changes
       M = M + 1
       CNT = CNT + 1
    END DO
    OUT(1) = CNT
    OUT(2) = I                     ! N + 1 after normal completion
    CNT = 0
    DO I = N, 1, -2
       CNT = CNT + 1
    END DO
    OUT(3) = CNT
    OUT(4) = I
    CNT = 0
    DO I = 5, 1                    ! zero trips
       CNT = CNT + 1
    END DO
    OUT(5) = CNT
    CNT = 0
    OUTER: DO I = 1, 4
       DO J = 1, 4
          IF (J == 3) CYCLE OUTER
          IF (I == 3) EXIT OUTER
          CNT = CNT + 1
       END DO
    END DO OUTER
    OUT(6) = CNT                   ! (1,1) (1,2) (2,1) (2,2) = 4
    OUT(7) = I                     ! 3
    CNT = 0
    DO I = 1, 10
       IF (I > 3) EXIT
       CNT = CNT + I
    END DO
    OUT(8) = CNT
  END SUBROUTINE DO_LOOPS

  SUBROUTINE BY_REFERENCE (X, OUT)
    REAL, INTENT(INOUT) :: X
    REAL, DIMENSION(6), INTENT(OUT) :: OUT
    REAL :: Y
    REAL, DIMENSION(-1:2) :: V
    V = 0.
    Y = 1.
    CALL BUMP (Y, 2.*X + 1., V(0), V(1:2))     ! variable, expression, element, section
    OUT(1) = Y
    OUT(2) = V(0)
    OUT(3) = V(1)
    OUT(4) = V(2)
    CALL BUMP (X, 0.5, SHARED_SCALAR, SHARED_VEC)   ! dummy and module variables
    OUT(5) = SHARED_SCALAR
    OUT(6) = SHARED_VEC(2)
  END SUBROUTINE BY_REFERENCE

  SUBROUTINE BUMP (A, INC, E, S)
    REAL, INTENT(INOUT) :: A
    REAL, INTENT(IN)    :: INC
    REAL, INTENT(OUT)   :: E
    REAL, DIMENSION(2), INTENT(OUT) :: S
    A = A + INC
    E = INC
    S(1) = A
    S(2) = -A
  END SUBROUTINE BUMP

  SUBROUTINE STATEMENT_FUNCTION_AND_HIDING (T, OUT)
    USE SEM_CONSTANTS
    REAL, INTENT(IN) :: T
    REAL, DIMENSION(4), INTENT(OUT) :: OUT
    REAL :: X, CLIP
    CLIP(X) = MIN( 50., MAX(-50.,(X-273.16)) )
    OUT(1) = CLIP(T)
    OUT(2) = CLIP(400.)
    OUT(3) = CICE                  ! the USEd module's, not the host module's
    OUT(4) = THIRD
  END SUBROUTINE STATEMENT_FUNCTION_AND_HIDING

  SUBROUTINE HOST_CONSTANT (OUT)
    REAL, DIMENSION(1), INTENT(OUT) :: OUT
    OUT(1) = CICE                  ! the host module's
  END SUBROUTINE HOST_CONSTANT

  SUBROUTINE SAVED_AND_DATA (OUT)
    REAL, DIMENSION(4), INTENT(OUT) :: OUT
    INTEGER :: NCALL
    REAL, DIMENSION(3) :: DZMIN
    DATA NCALL /0/
    SAVE NCALL
    DATA DZMIN /0.025, 0.025, 0.1/
    NCALL = NCALL + 1
    OUT(1) = NCALL
    OUT(2) = DZMIN(3)
    OUT(3) = TAB2(2,2)             ! module DATA with implied DO
    OUT(4) = TAB2(3,1)
  END SUBROUTINE SAVED_AND_DATA

  SUBROUTINE MASKS (N, A, L)
    INTEGER, INTENT(IN) :: N
    REAL, DIMENSION(N,2), INTENT(IN) :: A
    INTEGER, DIMENSION(N,2), INTENT(OUT) :: L
    WHERE (A-1.5.LT.0..AND.A.NE.-1.)
       L = 1
    ELSEWHERE
       L = -1
    ENDWHERE
  END SUBROUTINE MASKS

  SUBROUTINE MINMAX_NAN (X, OUT)
    REAL, INTENT(IN) :: X               ! a NaN
    REAL, DIMENSION(6), INTENT(OUT) :: OUT
    OUT(1) = MAX(X, 1.E-4)              ! gfortran: 1.E-4
    OUT(2) = MAX(1.E-4, X)              ! 1.E-4
    OUT(3) = MIN(X, 2.)
    OUT(4) = MAX(1., 2., 3.)
    OUT(5) = MIN(4, 2, 3)
    OUT(6) = SIGN(2., -0.)
  END SUBROUTINE MINMAX_NAN

  SUBROUTINE POWERS (X, N, OUT)
    REAL, INTENT(IN) :: X
    INTEGER, INTENT(IN) :: N
    REAL, DIMENSION(6), INTENT(OUT) :: OUT
    OUT(1) = X**5                       ! x*(x2*x2): libgcc's __powisf2 order
    OUT(2) = X**N
    OUT(3) = X**2.                      ! libm powf
    OUT(4) = 2**N                       ! integer
    OUT(5) = X**(-3)
    OUT(6) = X**0.5
  END SUBROUTINE POWERS

  SUBROUTINE OPTIONAL_ARGS (A, OUT, B, V)
    REAL, INTENT(IN) :: A
    REAL, DIMENSION(3), INTENT(OUT) :: OUT
    REAL, INTENT(IN), OPTIONAL :: B
    REAL, DIMENSION(2), INTENT(IN), OPTIONAL :: V
    OUT = 0.
    OUT(1) = A
    IF (PRESENT(B)) OUT(2) = B
    IF (PRESENT(V)) OUT(3) = V(2)
  END SUBROUTINE OPTIONAL_ARGS

  SUBROUTINE WITH_INTERNAL (X, OUT)
    REAL, INTENT(IN) :: X
    REAL, DIMENSION(2), INTENT(OUT) :: OUT
    REAL :: SCALE, Y
    SCALE = 3.
    CALL TIMES (X, Y)
    OUT(1) = Y
    SCALE = 4.
    CALL TIMES (X, Y)
    OUT(2) = Y
  CONTAINS
    SUBROUTINE TIMES (A, B)
      REAL, INTENT(IN)  :: A
      REAL, INTENT(OUT) :: B
      B = A * SCALE                    ! host association
    END SUBROUTINE TIMES
  END SUBROUTINE WITH_INTERNAL

  SUBROUTINE GOTOS (COSZ, OUT)
    REAL, INTENT(IN) :: COSZ
    REAL, DIMENSION(2), INTENT(OUT) :: OUT
    INTEGER :: NLOG, KCOUNT
    OUT = 0.
    IF (COSZ <= 0) GOTO 100
    OUT(1) = 1.
100 CONTINUE
    NLOG = 0
    KCOUNT = 0
1001 CONTINUE
    IF (.NOT.( (NLOG < 10) .AND. (KCOUNT == 0)))   goto 1002
    NLOG = NLOG + 1
    IF (NLOG == 4) KCOUNT = 1
    goto 1001
1002 CONTINUE
    OUT(2) = NLOG
  END SUBROUTINE GOTOS

  SUBROUTINE FATAL (N)
    INTEGER, INTENT(IN) :: N
    CHARACTER(len=256) :: message
    IF (N > 3) THEN
       WRITE(message,*) 'N = ', N
       call wrf_message(trim(message))
       call wrf_error_fatal ("too many " // "things")
    END IF
  END SUBROUTINE FATAL

  SUBROUTINE ARRAYS_3D (ims, ime, kms, kme, jms, jme, I, J, A, COL)
    INTEGER, INTENT(IN) :: ims, ime, kms, kme, jms, jme, I, J
    REAL, DIMENSION(ims:ime, kms:kme, jms:jme), INTENT(INOUT) :: A
    REAL, DIMENSION(kms:kme), INTENT(OUT) :: COL
    COL(kms:kme) = A(I, kms:kme, J)
    A(I, kms:kme, J) = 1.0
    A(I, kms, J) = COL(kme) / (COL(kms) + COL(kme))
  END SUBROUTINE ARRAYS_3D

END MODULE SEM_ROUTINES
