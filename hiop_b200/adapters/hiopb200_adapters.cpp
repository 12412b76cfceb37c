// Implementation of the HiOp-side adapter classes over the C-ABI (include/hiopb200.h).
#include "hiopb200_hooks.hpp"
#include "hiopLinSolverSymDenseB200.hpp"
#include "hiopKKTLinSysLowRankB200.hpp"
#include "hiopLinSolverSymDenseLapack.hpp"
#include "hiopNlpFormulation.hpp"
#include "hiopIterate.hpp"
#include "hiopResidual.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>

namespace hiop
{
namespace
{
hb_ctx* shared_ctx()
{
  static hb_ctx* ctx = nullptr;
  if(!ctx) {
    const char* dev = getenv("HIOP_B200_DEVICE");
    if(hb_ctx_create(dev ? atoi(dev) : 0, &ctx) != HB_OK) {
      fprintf(stderr, "hiop-b200: %s\n", hb_last_error());
      exit(EXIT_FAILURE); // like the reference's hard failures on missing back-ends (hiopNlpFormulation.cpp:360-363)
    }
  }
  return ctx;
}
void must(int rc, const char* what)
{
  if(rc != HB_OK) {
    fprintf(stderr, "hiop-b200: %s failed: %s\n", what, hb_last_error());
    exit(EXIT_FAILURE);
  }
}
} // namespace

bool hiop_b200_enabled()
{
  const char* e = getenv("HIOP_B200");
  return e && atoi(e) != 0;
}

hiopKKTLinSysLowRank* hiop_b200_new_lowrank_kkt(hiopNlpFormulation* nlp)
{
  if(hiop_b200_enabled()) return new hiopKKTLinSysLowRankB200(nlp);
  return new hiopKKTLinSysLowRank(nlp);
}

hiopLinSolverSymDense* hiop_b200_new_symdense_solver(int n, hiopNlpFormulation* nlp, bool safe_mode)
{
  if(hiop_b200_enabled()) {
    const char* force = getenv("HIOP_B200_LINSOL"); // "bk" | "nopiv": overrides the safe_mode choice
    int mode = safe_mode ? HB_FACT_BUNCH_KAUFMAN : HB_FACT_NOPIV;
    if(force && !strcmp(force, "bk")) mode = HB_FACT_BUNCH_KAUFMAN;
    if(force && !strcmp(force, "nopiv")) mode = HB_FACT_NOPIV;
    return new hiopLinSolverSymDenseB200(n, nlp, mode);
  }
  return new hiopLinSolverSymDenseLapack(n, nlp);
}

// ---------------------------------------------------------------------------------------------------------
hiopLinSolverSymDenseB200::hiopLinSolverSymDenseB200(int n, hiopNlpFormulation* nlp, int mode)
  : hiopLinSolverSymDense(n, nlp), ctx_(shared_ctx()), h_(nullptr), mode_(mode)
{
  must(hb_symdense_create(ctx_, n, &h_), "hb_symdense_create");
}
hiopLinSolverSymDenseB200::~hiopLinSolverSymDenseB200() { hb_symdense_destroy(h_); }

int hiopLinSolverSymDenseB200::matrixChanged()
{
  nlp_->runStats.linsolv.tmFactTime.start();
  // M_ lives in host memory (mem_space=default): upload + factorize (the MAGMA twin does the same H2D per factorization,
  // hiopLinSolverSymDenseMagma.cpp:139-146)
  const int ret = hb_symdense_matrix_changed_host(h_, M_->local_data(), mode_);
  nlp_->runStats.linsolv.tmFactTime.stop();
  if(ret < -1) {
    nlp_->log->printf(hovError, "hiopLinSolverSymDenseB200: %s\n", hb_last_error());
    return -1;
  }
  return ret;
}

bool hiopLinSolverSymDenseB200::solve(hiopVector& x)
{
  nlp_->runStats.linsolv.tmTriuSolves.start();
  const int rc = hb_symdense_solve_host(h_, x.local_data(), 1);
  nlp_->runStats.linsolv.tmTriuSolves.stop();
  return rc == 1;
}

// ---------------------------------------------------------------------------------------------------------
hiopKKTLinSysLowRankB200::hiopKKTLinSysLowRankB200(hiopNlpFormulation* nlp)
  : hiopKKTLinSysLowRank(nlp), ctx_(shared_ctx()), h_(nullptr), dJ_(nullptr), dSt_(nullptr), dYt_(nullptr)
{
  for(int i = 0; i < 12; i++) dres_[i] = ddir_[i] = nullptr;
  const char* ir = getenv("HIOP_B200_IR");
  ir_on_device_ = !(ir && !strcmp(ir, "host"));
  n_ = nlp_->n_local();
  meq_ = nlp_->m_eq();
  mineq_ = nlp_->m_ineq();
  lmax_ = nlp_->options->GetInteger("secant_memory_len");
  must(hb_lowrank_create(ctx_, n_, meq_, mineq_, lmax_ > 0 ? lmax_ : 1, &h_), "hb_lowrank_create");
  const size_t m = (size_t)meq_ + mineq_;
  must(hb_malloc(ctx_, sizeof(double) * m * n_, (void**)&dJ_), "hb_malloc(J)");
  must(hb_malloc(ctx_, sizeof(double) * (size_t)(lmax_ > 0 ? lmax_ : 1) * n_, (void**)&dSt_), "hb_malloc(S)");
  must(hb_malloc(ctx_, sizeof(double) * (size_t)(lmax_ > 0 ? lmax_ : 1) * n_, (void**)&dYt_), "hb_malloc(Y)");
  const size_t psz[4] = {(size_t)n_, (size_t)n_, (size_t)mineq_, (size_t)mineq_};
  for(int i = 0; i < 4; i++) must(hb_malloc(ctx_, sizeof(double) * psz[i], (void**)&dpat_[i]), "hb_malloc(pattern)");
  const size_t isz[8] = {(size_t)n_, (size_t)n_, (size_t)n_, (size_t)n_, (size_t)mineq_, (size_t)mineq_, (size_t)mineq_, (size_t)mineq_};
  for(int i = 0; i < 8; i++) must(hb_malloc(ctx_, sizeof(double) * isz[i], (void**)&dit_[i]), "hb_malloc(iterate)");
  const size_t rsz[3] = {(size_t)n_, (size_t)meq_, (size_t)mineq_};
  for(int i = 0; i < 3; i++) {
    must(hb_malloc(ctx_, sizeof(double) * rsz[i], (void**)&drhs_[i]), "hb_malloc(rhs)");
    must(hb_malloc(ctx_, sizeof(double) * rsz[i], (void**)&dsol_[i]), "hb_malloc(sol)");
  }
  // patterns are fixed for the lifetime of the formulation
  upload(dpat_[0], nlp_->get_ixl().local_data_const(), n_);
  upload(dpat_[1], nlp_->get_ixu().local_data_const(), n_);
  upload(dpat_[2], nlp_->get_idl().local_data_const(), mineq_);
  upload(dpat_[3], nlp_->get_idu().local_data_const(), mineq_);
  must(hb_lowrank_set_patterns(h_, dpat_[0], dpat_[1], dpat_[2], dpat_[3]), "hb_lowrank_set_patterns");
}

hiopKKTLinSysLowRankB200::~hiopKKTLinSysLowRankB200()
{
  hb_lowrank_destroy(h_);
  hb_free(ctx_, dJ_); hb_free(ctx_, dSt_); hb_free(ctx_, dYt_);
  for(auto* p : dpat_) hb_free(ctx_, p);
  for(auto* p : dit_) hb_free(ctx_, p);
  for(auto* p : drhs_) hb_free(ctx_, p);
  for(auto* p : dsol_) hb_free(ctx_, p);
  for(auto* p : dres_) if(p) hb_free(ctx_, p);
  for(auto* p : ddir_) if(p) hb_free(ctx_, p);
}

bool hiopKKTLinSysLowRankB200::upload(double* dst, const double* src, size_t count)
{
  if(count == 0) return true;
  must(hb_memcpy_h2d(ctx_, dst, src, sizeof(double) * count), "hb_memcpy_h2d");
  return true;
}

bool hiopKKTLinSysLowRankB200::update(const hiopIterate* iter, const hiopVector* grad_f, const hiopMatrixDense* Jac_c,
                                      const hiopMatrixDense* Jac_d, hiopHessianLowRank* Hess)
{
  // host bookkeeping of the reference (Dx_, Dd_inv_, DhInv are still read by the inherited computeDirections and by the
  // full-KKT operator of the outer BiCGStab refinement, hiopKKTLinSys.cpp:1619-1733)
  if(!hiopKKTLinSysLowRank::update(iter, grad_f, Jac_c, Jac_d, Hess)) return false;
  nlp_->runStats.tmSolverInternal.start();
  // Jacobian [Jc;Jd] -> device (the user callbacks write host memory, hiopNlpFormulation.cpp:1499-1533)
  upload(dJ_, Jac_c->local_data_const(), (size_t)meq_ * n_);
  upload(dJ_ + (size_t)meq_ * n_, Jac_d->local_data_const(), (size_t)mineq_ * n_);
  must(hb_lowrank_set_jacobian(h_, dJ_, dJ_ + (size_t)meq_ * n_), "hb_lowrank_set_jacobian");
  // secant memory as hiopHessianLowRank::update left it (hiopHessianLowRank.cpp:262-388)
  const int l = Hess->St_->m();
  if(l > 0) {
    upload(dSt_, Hess->St_->local_data_const(), (size_t)l * n_);
    upload(dYt_, Hess->Yt_->local_data_const(), (size_t)l * n_);
  }
  must(hb_lowrank_set_secant(h_, l, Hess->sigma, dSt_, dYt_, l ? Hess->L_->local_data_const() : nullptr,
                             l ? Hess->D_->local_data_const() : nullptr),
       "hb_lowrank_set_secant");
  const hiopVector* blocks[8] = {iter->zl, iter->sxl, iter->zu, iter->sxu, iter->vl, iter->sdl, iter->vu, iter->sdu};
  for(int i = 0; i < 8; i++) upload(dit_[i], blocks[i]->local_data_const(), blocks[i]->get_size());
  must(hb_lowrank_update(h_, dit_[0], dit_[1], dit_[2], dit_[3], dit_[4], dit_[5], dit_[6], dit_[7]), "hb_lowrank_update");
  // N and its factor depend only on the state set above: condense ONCE per update(); every preconditioner apply of the
  // outer BiCGStab then reuses it (the reference rebuilds and refactorizes N on each solveCompressed call).
  const int rc = hb_lowrank_condense(h_);
  nlp_->runStats.tmSolverInternal.stop();
  if(rc != HB_OK) {
    nlp_->log->printf(hovError, "hiopKKTLinSysLowRankB200::update: %s\n", hb_last_error());
    return false;
  }
  return true;
}

bool hiopKKTLinSysLowRankB200::solveCompressed(hiopVector& rx, hiopVector& ryc, hiopVector& ryd, hiopVector& dx, hiopVector& dyc,
                                               hiopVector& dyd)
{
  upload(drhs_[0], rx.local_data_const(), n_);
  upload(drhs_[1], ryc.local_data_const(), meq_);
  upload(drhs_[2], ryd.local_data_const(), mineq_);
  const int rc = hb_lowrank_solve_compressed(h_, drhs_[0], drhs_[1], drhs_[2], dsol_[0], dsol_[1], dsol_[2]);
  if(rc != HB_OK) {
    nlp_->log->printf(hovError, "hiopKKTLinSysLowRankB200::solveCompressed: %s\n", hb_last_error());
    return false;
  }
  if(n_) must(hb_memcpy_d2h(ctx_, dx.local_data(), dsol_[0], sizeof(double) * n_), "d2h");
  if(meq_) must(hb_memcpy_d2h(ctx_, dyc.local_data(), dsol_[1], sizeof(double) * meq_), "d2h");
  if(mineq_) must(hb_memcpy_d2h(ctx_, dyd.local_data(), dsol_[2], sizeof(double) * mineq_), "d2h");
  // like the reference, rx is overwritten with rx - J^T [dyc;dyd] (hiopKKTLinSys.cpp:1178)
  if(n_) must(hb_memcpy_d2h(ctx_, rx.local_data(), drhs_[0], sizeof(double) * n_), "d2h");
  must(hb_ctx_sync(ctx_), "hb_ctx_sync");
  return true;
}

bool hiopKKTLinSysLowRankB200::compute_directions_w_IR(const hiopResidual* resid, hiopIterate* dir)
{
  const int maxit = nlp_->options->GetInteger("ir_outer_maxit");
  if(!ir_on_device_ || maxit <= 0) return hiopKKTLinSys::compute_directions_w_IR(resid, dir);
  nlp_->runStats.tmSolverInternal.start();
  // compound order of hiopVectorCompoundPD (hiopVectorCompoundPD.cpp:228-255)
  const hiopVector* rb[12] = {resid->rx, resid->rd, resid->ryc, resid->ryd, resid->rxl, resid->rxu, resid->rdl, resid->rdu,
                              resid->rszl, resid->rszu, resid->rsvl, resid->rsvu};
  hiopVector* db[12] = {dir->x, dir->d, dir->yc, dir->yd, dir->sxl, dir->sxu, dir->sdl, dir->sdu, dir->zl, dir->zu, dir->vl, dir->vu};
  for(int i = 0; i < 12; i++) {
    const size_t sz = (size_t)rb[i]->get_size();
    if(!dres_[i]) {
      must(hb_malloc(ctx_, sizeof(double) * sz, (void**)&dres_[i]), "hb_malloc(residual)");
      must(hb_malloc(ctx_, sizeof(double) * sz, (void**)&ddir_[i]), "hb_malloc(direction)");
    }
    upload(dres_[i], rb[i]->local_data_const(), sz);
  }
  const double tol = std::min(mu_ * nlp_->options->GetNumeric("ir_outer_tol_factor"), nlp_->options->GetNumeric("ir_outer_tol_min"));
  double info[4] = {0, 0, 0, 0};
  const int rc = hb_lowrank_compute_directions_w_ir(h_, dres_, ddir_, tol, maxit, info);
  if(rc != HB_OK) {
    nlp_->log->printf(hovError, "hiopKKTLinSysLowRankB200::compute_directions_w_IR: %s\n", hb_last_error());
    nlp_->runStats.tmSolverInternal.stop();
    return false;
  }
  for(int i = 0; i < 12; i++) {
    const size_t sz = (size_t)db[i]->get_size();
    if(sz) must(hb_memcpy_d2h(ctx_, db[i]->local_data(), ddir_[i], sizeof(double) * sz), "d2h");
  }
  must(hb_ctx_sync(ctx_), "hb_ctx_sync");
  nlp_->runStats.kkt.nIterRefinInner += info[1];
  if(info[0] != 0.0)  // the step is accepted whatever BiCGStab reports (hiopKKTLinSys.cpp:950-953)
    nlp_->log->printf(hovWarning, "BiCGStab (device) did NOT converge: flag %d after %g iters, abs res %g, rel res %g\n", (int)info[0], info[1],
                      info[2], info[3]);
  else
    nlp_->log->printf(hovScalars, "BiCGStab (device) converged: actual normResid=%g relResid=%g iter=%g\n", info[2], info[3], info[1]);
  nlp_->runStats.tmSolverInternal.stop();
  return true;
}

} // namespace hiop
