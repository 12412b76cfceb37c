// Factory hooks a HiOp maintainer adds at the three places where the reference hard-codes its KKT / linear-solver / quasi-Newton
// classes:
//   src/Optimization/hiopAlgFilterIPM.cpp:1050      hiopKKTLinSysLowRank* kkt = new hiopKKTLinSysLowRank(nlp);
//   src/Optimization/hiopKKTLinSysMDS.cpp:437-478   linSys_ = new hiopLinSolverSymDenseLapack(n, nlp_);
//   src/Optimization/hiopNlpFormulation.cpp:1639    return new hiopHessianLowRank(this, secant_memory_len);
// Selection is by the environment variable HIOP_B200 (unset/0 -> the reference classes, 1 -> the B200 engine), so the
// options parser (src/Utils/hiopOptions.cpp) is not forked.
#pragma once
namespace hiop
{
class hiopNlpFormulation;
class hiopKKTLinSysLowRank;
class hiopLinSolverSymDense;
class hiopNlpDenseConstraints;
class hiopMatrix;

hiopKKTLinSysLowRank* hiop_b200_new_lowrank_kkt(hiopNlpFormulation* nlp);
hiopLinSolverSymDense* hiop_b200_new_symdense_solver(int n, hiopNlpFormulation* nlp, const bool* safe_mode);
/// hiopHessianLowRankB200 when the engine is selected (its update() runs on the device under HIOP_B200_SECANT=device), else the reference class
hiopMatrix* hiop_b200_new_hessian_lowrank(hiopNlpDenseConstraints* nlp, int max_memory_length);
bool hiop_b200_enabled();
} // namespace hiop
