"""oracle/f90interp.py implements Fortran semantics, nothing about interpolation: these small Fortran programs
check the rules the golden vectors depend on (CPU only)."""
import numpy as np
import pytest

from oracle.f90interp import Cell, Interpreter

SRC = """
MODULE m
  IMPLICIT NONE
  TYPE :: node
    REAL :: p(2)
    INTEGER :: id
    TYPE(node), POINTER :: next => NULL()
  END TYPE node
  TYPE :: pair
    INTEGER :: a
    REAL :: b
  END TYPE pair
  CONTAINS

  SUBROUTINE loops(n, after_full, after_exit, trips_neg, zero_trip)
    INTEGER, INTENT(IN) :: n
    INTEGER, INTENT(OUT) :: after_full, after_exit, trips_neg, zero_trip
    INTEGER :: i
    DO i = 1, n
    END DO
    after_full = i            ! n + 1
    DO i = n, 2, -1
      IF (i == 4) EXIT
    END DO
    after_exit = i            ! 4
    trips_neg = 0
    DO i = -3, 3, 6           ! -3 and 3
      trips_neg = trips_neg + 1
    END DO
    zero_trip = 0
    DO i = 5, 4
      zero_trip = zero_trip + 1
    END DO
  END SUBROUTINE loops

  SUBROUTINE arith(q1, q2, q3, m1, m2, r1, r2, t)
    INTEGER, INTENT(OUT) :: q1, q2, q3, m1, m2
    REAL, INTENT(OUT) :: r1, r2
    INTEGER, INTENT(OUT) :: t
    REAL :: x
    q1 = 7 / 2                ! 3
    q2 = (-7) / 2             ! -3: truncation toward zero
    q3 = (5 + 1) / 2
    m1 = MOD(-7, 3)           ! -1: sign of the dividend
    m2 = MODULO(-7, 3)        ! 2: sign of the divisor
    r1 = 7 / 2                ! integer division first, then conversion: 3.0
    r2 = REAL(7) / 2          ! 3.5
    x = -2.7
    t = x                     ! real -> integer assignment truncates: -2
  END SUBROUTINE arith

  SUBROUTINE sections(a, b, cnt, mx)
    REAL, DIMENSION(:, :), INTENT(INOUT) :: a     ! (3, 4)
    REAL, DIMENSION(:), INTENT(INOUT) :: b        ! (5)
    INTEGER, INTENT(OUT) :: cnt
    REAL, INTENT(OUT) :: mx
    REAL :: tmp(3)
    tmp = a(:, 1)
    a(:, 1) = a(:, 4)         ! column swap through sections
    a(:, 4) = tmp
    a(2, 2:3) = 0.0
    b(2:5) = b(1:4)           ! overlapping shift: right-hand side is evaluated first
    cnt = COUNT(b < 2.5)
    mx = MAXVAL(a(3, :))
    CALL bump(a(:, 2))        ! a section is passed by reference
  END SUBROUTINE sections

  SUBROUTINE bump(v)
    REAL, DIMENSION(:), INTENT(INOUT) :: v
    v(1) = v(1) + 100.0
  END SUBROUTINE bump

  RECURSIVE SUBROUTINE push(head, id, depth)
    TYPE(node), POINTER :: head
    INTEGER, INTENT(IN) :: id, depth
    TYPE(node), POINTER :: fresh
    IF (depth > 0) THEN
      CALL push(head, id, depth - 1)   ! recursion with a pointer dummy
      RETURN
    END IF
    ALLOCATE(fresh)
    fresh%id = id
    fresh%p = [REAL(id), REAL(id) * 0.5]
    fresh%next => head
    head => fresh                      ! re-points the caller's pointer
  END SUBROUTINE push

  SUBROUTINE walk(head, total, length)
    TYPE(node), POINTER :: head
    REAL, INTENT(OUT) :: total
    INTEGER, INTENT(OUT) :: length
    TYPE(node), POINTER :: cur
    total = 0.0
    length = 0
    cur => head
    DO WHILE (ASSOCIATED(cur))
      total = total + cur%p(1) + cur%p(2)
      length = length + 1
      cur => cur%next
    END DO
  END SUBROUTINE walk

  SUBROUTINE structs(n, s)
    INTEGER, INTENT(IN) :: n
    INTEGER, INTENT(OUT) :: s
    TYPE(pair), DIMENSION(10) :: x, y
    INTEGER :: i
    DO i = 1, n
      x(i)%a = i
      x(i)%b = 0.25 * i
    END DO
    y(:n) = x(:n)              ! element-wise copy of derived types
    x(1)%a = 99                ! must not show through y
    s = 0
    DO i = 1, n
      s = s + y(i)%a
    END DO
  END SUBROUTINE structs

  PURE FUNCTION sq(u, v) RESULT(r)
    REAL, INTENT(IN) :: u(2), v(2)
    REAL :: r
    r = (u(1) - v(1)) ** 2 + (u(2) - v(2)) ** 2
  END FUNCTION sq

  SUBROUTINE kinds(big, small, s3, acc)
    REAL, INTENT(OUT) :: big, small, s3, acc
    REAL :: w(4)
    INTEGER :: i
    big = HUGE(1.0)
    small = TINY(1.0) * 1e4
    s3 = sq([1.0, 2.0], [4.0, 6.0])
    w = (/(0.1 * i, i = 1, 4)/)
    acc = SUM(w)               ! sequential sum in the default real kind
  END SUBROUTINE kinds
END MODULE m
"""


@pytest.fixture(scope="module", params=[np.float64, np.float32])
def itp(request):
    return Interpreter([SRC], request.param)


def outs(n, v=0):
    return [Cell(v) for _ in range(n)]


def test_do_loop_variable_semantics(itp):
    o = outs(4)
    itp.call("loops", 7, *o)
    assert [c.get() for c in o] == [8, 4, 2, 0]


def test_integer_and_mixed_arithmetic(itp):
    o = outs(8)
    itp.call("arith", *o)
    q1, q2, q3, m1, m2, r1, r2, t = [c.get() for c in o]
    assert (q1, q2, q3, m1, m2, t) == (3, -3, 3, -1, 2, -2)
    assert r1 == 3.0 and r2 == 3.5 and isinstance(r2, itp.real)


def test_sections_are_views_and_rhs_is_evaluated_first(itp):
    R = itp.real
    a = np.asfortranarray(np.arange(12, dtype=R).reshape(3, 4))
    b = np.array([1, 2, 3, 4, 5], dtype=R)
    cnt, mx = Cell(0), Cell(R(0))
    a0 = a.copy()
    itp.call("sections", a, b, cnt, mx)
    assert np.array_equal(a[:, 0], a0[:, 3]) and np.array_equal(a[:, 3], a0[:, 0])
    assert a[1, 1] == 100.0 - 100.0 and a[1, 2] == 0.0 and a[0, 1] == a0[0, 1] + 100.0
    assert np.array_equal(b, [1, 1, 2, 3, 4]) and cnt.get() == 3 and mx.get() == a0[2].max()


def test_pointers_recursion_and_derived_types(itp):
    head = Cell(None)
    head.decl = itp.prog.procs["push"].decls["head"]
    for i in (1, 2, 3):
        itp.call("push", head, i, 2)
    total, length = Cell(itp.real(0)), Cell(0)
    itp.call("walk", head, total, length)
    assert length.get() == 3 and total.get() == 1.5 * 6 and head.get().f["id"] == 3
    s = Cell(0)
    itp.call("structs", 4, s)
    assert s.get() == 10


def test_default_real_kind_rules(itp):
    R = itp.real
    o = outs(4, R(0))
    itp.call("kinds", *o)
    big, small, s3, acc = [c.get() for c in o]
    assert big == np.finfo(R).max and small == R(np.finfo(R).tiny) * R(1e4) and s3 == R(25.0)
    w = [R(0.1) * R(i) for i in (1, 2, 3, 4)]
    assert acc == ((w[0] + w[1]) + w[2]) + w[3] and isinstance(acc, R)
