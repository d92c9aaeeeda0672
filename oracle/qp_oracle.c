/*
 * ORACLE (test infrastructure, NOT product code).
 * Instantiates the CPU restatement of /root/reference/src/qp.cpp for double and float,
 * mirroring "template class QPSolver<double>; template class QPSolver<float>;"
 * (src/qp.cpp:385-386). See qp_oracle_impl.h for the per-function citations and the
 * parity-pinning statement.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* include/solvers/qp.hpp:70 */
enum { ORC_SOLVED = 0, ORC_MAX_ITER_EXCEEDED = 1, ORC_UNSOLVED = 2, ORC_NUMERICAL_ISSUES = 3, ORC_UNINITIALIZED = 4 };
/* include/solvers/qp.hpp:134 */
enum { ORC_INEQUALITY_CONSTRAINT = 0, ORC_EQUALITY_CONSTRAINT = 1, ORC_LOOSE_BOUNDS = 2 };

#define SCALAR double
#define SUFFIX _f64
#define ORC_FABS fabs
#define ORC_FMAX fmax
#define ORC_FMIN fmin
#define ORC_RHO_EST(rho0, arg) ((rho0) * sqrt(arg)) /* qp.cpp:338 */
#define ORC_EPS DBL_EPSILON /* DIV_BY_ZERO_REGUL, qp.hpp:141 */
#define ORC_MINPOS DBL_MIN
#include "qp_oracle_impl.h"
#undef SCALAR
#undef SUFFIX
#undef ORC_FABS
#undef ORC_FMAX
#undef ORC_FMIN
#undef ORC_RHO_EST
#undef ORC_EPS
#undef ORC_MINPOS

#define SCALAR float
#define SUFFIX _f32
#define ORC_FABS fabsf
/* the reference calls the double overloads fmax/fmin/sqrt on float arguments (src/qp.cpp:131, :338) */
#define ORC_FMAX(a, b) ((float)fmax((double)(a), (double)(b)))
#define ORC_FMIN(a, b) ((float)fmin((double)(a), (double)(b)))
/* qp.cpp:338 `Scalar rho_new = rho0 * sqrt(...)`: with <cmath> only, unqualified sqrt(float) is ::sqrt(double) and returns a double, so
 * the PRODUCT is formed in double and rounded to float once, on assignment (verified against oracle/_ref, the reference's own code) */
#define ORC_RHO_EST(rho0, arg) ((float)((double)(rho0) * sqrt((double)(arg))))
#define ORC_EPS FLT_EPSILON
#define ORC_MINPOS FLT_MIN
#include "qp_oracle_impl.h"

int oracle_num_procs(void) {
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}
