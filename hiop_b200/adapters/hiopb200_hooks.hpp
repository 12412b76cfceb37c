// Factory hooks a HiOp maintainer adds at the two places where the reference hard-codes its KKT / linear-solver
// classes:
//   src/Optimization/hiopAlgFilterIPM.cpp:1050      hiopKKTLinSysLowRank* kkt = new hiopKKTLinSysLowRank(nlp);
//   src/Optimization/hiopKKTLinSysMDS.cpp:437-478   linSys_ = new hiopLinSolverSymDenseLapack(n, nlp_);
// Selection is by the environment variable HIOP_B200 (unset/0 -> the reference classes, 1 -> the B200 engine), so the
// options parser (src/Utils/hiopOptions.cpp) is not forked.
#pragma once
namespace hiop
{
class hiopNlpFormulation;
class hiopKKTLinSysLowRank;
class hiopLinSolverSymDense;

hiopKKTLinSysLowRank* hiop_b200_new_lowrank_kkt(hiopNlpFormulation* nlp);
hiopLinSolverSymDense* hiop_b200_new_symdense_solver(int n, hiopNlpFormulation* nlp, const bool* safe_mode);
bool hiop_b200_enabled();
} // namespace hiop
