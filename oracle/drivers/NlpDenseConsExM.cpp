// TEST INFRASTRUCTURE (built by oracle/Makefile against the reference library, never linked into libhiopb200.so).
//
// Dense-constraints NLP with a PARAMETRISED number of constraints, the generalisation SURVEY.md section 0 item 4 asks for
// (the bundled NlpDenseConsEx1 / Ex2 have m = 1 / m <= 4, so the int8-slice condensation -- the default for m + 2l >= 64 -- never ran
// inside an interior-point loop). Same interface the bundled drivers implement (hiopInterfaceDenseConstraints,
// src/Interface/hiopInterface.hpp:517-570; compare src/Drivers/Dense/NlpDenseConsEx1.hpp:136-137 for the class being generalised):
//
//    min   sum_i 0.25 (x_i - a_i)^4 + 0.5 w_i x_i^2
//    s.t.  (A x)_j  = b_j            j < m_eq
//          b_j - 1 <= (A x)_j <= (b_j + 1 | +inf)   m_eq <= j < m
//          0 <= x_i ( <= 2 on every 5th variable)
//
// A (m x n, dense) = smooth rows (low-frequency cosines, all entries nonzero) scaled by 1/sqrt(n); b = A x_ref with x_ref interior,
// so the problem is feasible. The constraints are linear: "-linear" makes get_prob_info say so (HiOp then evaluates the Jacobian
// once, hiopNlpFormulation.cpp:1548-1551) -- without it the Jacobian is re-evaluated (with identical values) every iteration,
// like the bundled drivers do.
//
// Usage: exM.exe n m [-linear] [-selfcheck]      prints HiOp's iteration table (the drop-in tests diff it) and the final objective.
#include "hiopInterface.hpp"
#include "hiopNlpFormulation.hpp"
#include "hiopAlgFilterIPM.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace hiop;

class DenseConsExM : public hiopInterfaceDenseConstraints
{
public:
  DenseConsExM(int n, int m, bool declare_linear) : n_(n), m_(m), meq_(m - m / 2), linear_(declare_linear), A_((size_t)m * n), b_(m), a_(n), w_(n)
  {
    const double pi = 3.14159265358979323846;
    for(int j = 0; j < m_; j++)
      for(int i = 0; i < n_; i++) {
        const double t = (i + 0.5) / n_;
        A_[(size_t)j * n_ + i] = (j == 0 ? 1.0 : std::cos(pi * j * t) + 0.3 * std::sin(2.0 * pi * (j % 7 + 1) * t)) / std::sqrt((double)n_);
      }
    for(int i = 0; i < n_; i++) {
      const double t = (i + 0.5) / n_;
      a_[i] = 1.0 + 0.5 * std::sin(6.0 * pi * t);
      w_[i] = 0.1 + 0.05 * std::cos(4.0 * pi * t);
    }
    for(int j = 0; j < m_; j++) {
      double s = 0.0;
      for(int i = 0; i < n_; i++) s += A_[(size_t)j * n_ + i] * x_ref(i);
      b_[j] = s;
    }
  }
  double x_ref(int i) const { return 0.6 + 0.3 * std::cos(10.0 * (i + 0.5) / n_); }

  bool get_prob_sizes(size_type& n, size_type& m) { n = n_; m = m_; return true; }
  bool get_prob_info(NonlinearityType& type) { type = linear_ ? hiopLinear : hiopNonlinear; return true; }
  bool get_vars_info(const size_type& n, double* xlow, double* xupp, NonlinearityType* type)
  {
    for(int i = 0; i < n_; i++) {
      xlow[i] = 0.0;
      xupp[i] = (i % 5 == 0) ? 2.0 : 1e20;
      type[i] = hiopNonlinear;
    }
    return true;
  }
  bool get_cons_info(const size_type& m, double* clow, double* cupp, NonlinearityType* type)
  {
    for(int j = 0; j < m_; j++) {
      if(j < meq_) { clow[j] = cupp[j] = b_[j]; }
      else { clow[j] = b_[j] - 1.0; cupp[j] = ((j - meq_) % 3 == 0) ? b_[j] + 1.0 : 1e20; }
      type[j] = hiopLinear;
    }
    return true;
  }
  bool eval_f(const size_type& n, const double* x, bool new_x, double& f)
  {
    f = 0.0;
    for(int i = 0; i < n_; i++) {
      const double d = x[i] - a_[i];
      f += 0.25 * d * d * d * d + 0.5 * w_[i] * x[i] * x[i];
    }
    return true;
  }
  bool eval_grad_f(const size_type& n, const double* x, bool new_x, double* g)
  {
    for(int i = 0; i < n_; i++) {
      const double d = x[i] - a_[i];
      g[i] = d * d * d + w_[i] * x[i];
    }
    return true;
  }
  bool eval_cons(const size_type& n, const size_type& m, const size_type& num_cons, const index_type* idx_cons, const double* x, bool new_x, double* cons)
  {
    for(int k = 0; k < (int)num_cons; k++) {
      const double* row = &A_[(size_t)idx_cons[k] * n_];
      double s = 0.0;
      for(int i = 0; i < n_; i++) s += row[i] * x[i];
      cons[k] = s;
    }
    return true;
  }
  bool eval_Jac_cons(const size_type& n, const size_type& m, const size_type& num_cons, const index_type* idx_cons, const double* x, bool new_x, double* Jac)
  {
    for(int k = 0; k < (int)num_cons; k++) memcpy(Jac + (size_t)k * n_, &A_[(size_t)idx_cons[k] * n_], sizeof(double) * n_);
    return true;
  }
  bool get_starting_point(const size_type& n, double* x0)
  {
    for(int i = 0; i < n_; i++) x0[i] = 1.0;
    return true;
  }

private:
  int n_, m_, meq_;
  bool linear_;
  std::vector<double> A_, b_, a_, w_;
};

int main(int argc, char** argv)
{
  if(argc < 3) {
    printf("usage: %s n m [-linear] [-selfcheck]\n", argv[0]);
    return 1;
  }
  const int n = atoi(argv[1]), m = atoi(argv[2]);
  bool linear = false;
  for(int a = 3; a < argc; a++)
    if(!strcmp(argv[a], "-linear")) linear = true;
  DenseConsExM problem(n, m, linear);
  hiopNlpDenseConstraints nlp(problem);
  nlp.options->SetStringValue("duals_update_type", "linear");
  nlp.options->SetStringValue("duals_init", "zero");
  nlp.options->SetStringValue("Hessian", "quasinewton_approx");
  nlp.options->SetStringValue("KKTLinsys", "xdycyd");
  nlp.options->SetStringValue("compute_mode", "cpu");
  nlp.options->SetIntegerValue("verbosity_level", 3);
  nlp.options->SetNumericValue("mu0", 1e-1);
  nlp.options->SetNumericValue("tolerance", 1e-6);
  nlp.options->SetIntegerValue("max_iter", 60);
  hiopAlgFilterIPM solver(&nlp);
  const hiopSolveStatus status = solver.run();
  const double obj = solver.getObjective();
  printf("ExM n=%d m=%d status=%d iterations=%d objective=%.12e\n", n, m, (int)status, solver.getNumIterations(), obj);
  return status < 0 ? 2 : 0;
}
