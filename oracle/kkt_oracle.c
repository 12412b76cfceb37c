/* TEST INFRASTRUCTURE ONLY -- a plain-C restatement of the reference's scalar hot loops, used as the parity
 * checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg. Never called by the product path.
 *
 * Every function follows the loop order of the reference function it cites so that results agree with the
 * compiled reference (oracle/_ref) to the last bit where the reference itself is deterministic.
 * Paths are relative to /root/reference/.
 */
#include <math.h>
#include <stddef.h>
#include <string.h>

/* W = beta*W + alpha * X*diag(d)*X^T, X is k x n row-major, W is k x k row-major, both triangles written.
 * src/Optimization/hiopHessianLowRank.cpp:1079-1116 (symmMatTimesDiagTimesMatTrans_local) */
void ko_symm_mat_diag_mat_trans(double beta, double* W, double alpha, const double* X, const double* d, int k, long n)
{
  for(int i = 0; i < k; i++) {
    const double* xi = X + (size_t)i * n;
    for(int j = i; j < k; j++) {
      const double* xj = X + (size_t)j * n;
      double acc = 0.0;
      for(long p = 0; p < n; p++) acc += xi[p] * d[p] * xj[p];
      W[(size_t)i * k + j] = W[(size_t)j * k + i] = beta * W[(size_t)i * k + j] + alpha * acc;
    }
  }
}

/* W = S*diag(d)*X^T, S is l x n, X is k x n, W is l x k (row stride ldw).
 * src/Optimization/hiopHessianLowRank.cpp:1119-1154 (matTimesDiagTimesMatTrans_local) */
void ko_mat_diag_mat_trans(double* W, int ldw, const double* S, const double* d, const double* X, int l, int k, long n)
{
  for(int i = 0; i < l; i++) {
    const double* Si = S + (size_t)i * n;
    for(int j = 0; j < k; j++) {
      const double* Xj = X + (size_t)j * n;
      double acc = 0.;
      for(long p = 0; p < n; p++) acc += Si[p] * d[p] * Xj[p];
      W[(size_t)i * ldw + j] = acc;
    }
  }
}

/* y += alpha*x/z where select==1.  src/LinAlg/hiopVectorPar.cpp:767-790 (alpha = +-1 special-cased there) */
void ko_axdzpy_w_pattern(long n, double* y, double alpha, const double* x, const double* z, const double* sel)
{
  if(alpha == 1.0) {
    for(long i = 0; i < n; i++) if(sel[i] == 1.0) y[i] += x[i] / z[i];
  } else if(alpha == -1.0) {
    for(long i = 0; i < n; i++) if(sel[i] == 1.0) y[i] -= x[i] / z[i];
  } else {
    for(long i = 0; i < n; i++) if(sel[i] == 1.0) y[i] += alpha * x[i] / z[i];
  }
}

/* y += alpha*x*z. src/LinAlg/hiopVectorPar.cpp:710-734 */
void ko_axzpy(long n, double* y, double alpha, const double* x, const double* z)
{
  if(alpha == 1.0) {
    for(long i = 0; i < n; i++) y[i] += x[i] * z[i];
  } else if(alpha == -1.0) {
    for(long i = 0; i < n; i++) y[i] -= x[i] * z[i];
  } else {
    for(long i = 0; i < n; i++) y[i] += alpha * x[i] * z[i];
  }
}

/* y = sel ? y/x : 0. src/LinAlg/hiopVectorPar.cpp:584-592 */
void ko_component_div_w_select(long n, double* y, const double* x, const double* sel)
{
  for(long i = 0; i < n; i++) {
    if(sel[i] == 0.0) y[i] = 0.0;
    else y[i] /= x[i];
  }
}

/* sum log(y_i) over select==1, Kahan-compensated. src/LinAlg/hiopVectorPar.cpp:863-881 */
double ko_log_barrier(long n, const double* y, const double* sel)
{
  double sum = 0.0, comp = 0.0;
  for(long i = 0; i < n; i++) {
    if(sel[i] != 0.0) {
      double logval = log(y[i]);
      logval -= comp;
      double aux = sum + logval;
      comp = (aux - sum) - logval;
      sum = aux;
    }
  }
  return sum;
}

/* y += alpha/x where select==1. src/LinAlg/hiopVectorPar.cpp:893-905 */
void ko_add_log_barrier_grad(long n, double* y, double alpha, const double* x, const double* sel)
{
  for(long i = 0; i < n; i++) if(sel[i] == 1.0) y[i] += alpha / x[i];
}

/* mu*kappa_d * sum y_i over ixl==1 && ixu==0. src/LinAlg/hiopVectorPar.cpp:907-925 */
double ko_linear_damping_term(long n, const double* y, const double* ixl, const double* ixu, double mu, double kappa_d)
{
  double term = 0.0;
  for(long i = 0; i < n; i++) if(ixl[i] == 1.0 && ixu[i] == 0.0) term += y[i];
  term *= mu;
  term *= kappa_d;
  return term;
}

/* y = alpha*y + ct*(ixl-ixu). src/LinAlg/hiopVectorPar.cpp:927-944 */
void ko_add_linear_damping_term(long n, double* y, const double* ixl, const double* ixu, double alpha, double ct)
{
  for(long i = 0; i < n; i++) y[i] = alpha * y[i] + ct * (ixl[i] - ixu[i]);
}

/* y = beta*y + alpha*A*x, A m x n row-major (the reference calls DGEMV 'T' on the column-major view).
 * src/LinAlg/hiopMatrixDenseRowMajor.cpp:436-492 */
void ko_times_vec(int m, long n, const double* A, double beta, double* y, double alpha, const double* x)
{
  for(int i = 0; i < m; i++) {
    const double* Ai = A + (size_t)i * n;
    double acc = 0.;
    for(long p = 0; p < n; p++) acc += Ai[p] * x[p];
    y[i] = beta * y[i] + alpha * acc;
  }
}

/* y = beta*y + alpha*A^T*x. src/LinAlg/hiopMatrixDenseRowMajor.cpp:494-528 */
void ko_trans_times_vec(int m, long n, const double* A, double beta, double* y, double alpha, const double* x)
{
  if(beta == 0.0) memset(y, 0, sizeof(double) * n);
  else if(beta != 1.0) for(long p = 0; p < n; p++) y[p] *= beta;
  for(int i = 0; i < m; i++) {
    const double* Ai = A + (size_t)i * n;
    const double ax = alpha * x[i];
    for(long p = 0; p < n; p++) y[p] += ax * Ai[p];
  }
}
