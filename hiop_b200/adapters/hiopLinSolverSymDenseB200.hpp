// hiopLinSolverSymDenseB200 -- drop-in hiopLinSolverSymDense (src/LinAlg/hiopLinSolver.hpp:78-128) backed by libhiopb200.so.
// Same contract as hiopLinSolverSymDenseLapack (src/LinAlg/hiopLinSolverSymDenseLapack.hpp:75-192) and the MAGMA twins
// (src/LinAlg/hiopLinSolverSymDenseMagma.cpp:120-270, 324-476): the KKT class fills the upper triangle of sysMatrix()
// (host memory when mem_space=default), matrixChanged() factorizes and returns #negative eigenvalues or -1,
// solve(x) overwrites the right-hand side.
#pragma once
#include "hiopLinSolver.hpp"
#include "hiopb200.h"

namespace hiop
{
class hiopLinSolverSymDenseB200 : public hiopLinSolverSymDense
{
public:
  /// safe_mode: pointer to the owning KKT object's safe-mode flag, read at every matrixChanged(): Bunch-Kaufman while it is set
  /// (MagmaBuKa role), LDL^T without pivoting otherwise (MagmaNopiv role). NULL = always Bunch-Kaufman.
  hiopLinSolverSymDenseB200(int n, hiopNlpFormulation* nlp, const bool* safe_mode);
  virtual ~hiopLinSolverSymDenseB200();
  int matrixChanged() override;
  bool solve(hiopVector& x) override;

private:
  hb_ctx* ctx_;
  hb_symdense* h_;
  const bool* safe_mode_;
  bool healthy_;
};
} // namespace hiop
