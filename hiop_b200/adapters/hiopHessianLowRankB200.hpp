// hiopHessianLowRankB200 -- a11 in the adapter: the secant memory of hiopHessianLowRank (src/Optimization/hiopHessianLowRank.cpp:262-388)
// kept in HBM. The reference's update() forms s and y on the host with four J^T products and two Jacobian copies per iteration (48 GB of
// host traffic at n = 1e6, m = 1000) and the KKT adapter then uploads S_t and Y_t again. With HIOP_B200_SECANT=device this class turns
// update() into a note "one more iterate arrived"; hiopKKTLinSysLowRankB200::update -- called by hiopAlgFilterIPMQuasiNewton::run with the
// same iterate, gradient and Jacobians on the very next line (src/Optimization/hiopAlgFilterIPM.cpp:1215-1216) -- hands them to
// hb_lowrank_secant_update, which owns S_t, Y_t, x_prev, grad_f_prev and J_prev on the device (hb_secant.cu) and mirrors only l, sigma, L and
// D back into this object.
//
// What stays valid on the host in that mode: l_curr, sigma, L_, D_. S_t / Y_t of this object are NOT maintained: the inherited solve(),
// timesVec() and symMatTimesInverseTimesMatTrans() must not be used (the KKT adapter does not; HIOP_B200_IR=host, whose host-side BiCGStab
// needs timesVec, therefore falls back to the host secant update).
#pragma once
#include "hiopHessianLowRank.hpp"

namespace hiop
{
class hiopHessianLowRankB200 : public hiopHessianLowRank
{
public:
  hiopHessianLowRankB200(hiopNlpDenseConstraints* nlp, int max_memory_length);
  virtual ~hiopHessianLowRankB200() {}

  bool update(const hiopIterate& x_curr, const hiopVector& grad_f_curr, const hiopMatrix& Jac_c_curr, const hiopMatrix& Jac_d_curr) override;

  bool device_mode() const { return device_mode_; }
  /// number of update() calls since the last take_pending(); the KKT adapter performs them on the device
  int take_pending() { const int p = pending_; pending_ = 0; return p; }
  /// gradient of the objective handed to the last update(): the KKT update() cannot supply it -- the reference's non-virtual wrapper
  /// forwards its own (never set) member instead of the argument (src/Optimization/hiopKKTLinSys.hpp:404)
  const hiopVector* pending_grad_f() const { return pending_grad_f_; }
  const hiopIterate* pending_iterate() const { return pending_it_; }
  /// B0 scaling and its update rule as read from the options by the base class (hiopHessianLowRank.cpp:118-136)
  double sigma0_value() const;
  int sigma_strategy_value() const;
  /// scalars of the device-side state mirrored into the host object
  void mirror(int l, double sigma_new, const double* L_host, const double* D_host);

private:
  bool device_mode_;
  int pending_;
  const hiopVector* pending_grad_f_ = nullptr;
  const hiopIterate* pending_it_ = nullptr;
};
} // namespace hiop
