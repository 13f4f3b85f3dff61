// Xyce-side adaptor: the BSIM4 device master behind Xyce's own plugin interface, evaluated by libxyce_b200.so.
//
// This is the reference-side binding of the C ABI (include/xyce_b200.h) for the BSIM4 device: a drop-in replacement of
// Xyce::Device::MOSFET_B4::Master (src/DeviceModelPKG/OpenModels/N_DEV_MOSFET_B4.h) created through the existing
// factory hook -- Traits::factory (N_DEV_MOSFET_B4.C:11688-11692) returns `new GpuMaster(...)` instead of
// `new Master(...)`.  DeviceMgr, the loaders, the time integrator and the nonlinear solver keep calling the same
// virtuals (Core/N_DEV_Device.h:312-453, Core/N_DEV_DeviceMaster.h:335-343):
//     updateState(sol, sta, sto) -> loadDAEVectors(sol, f, q, b, leadF, leadQ, junctionV) -> loadDAEMatrices(dFdx, dQdx)
// and get the same numbers (1e-12) from the GPU.
//
// This version serves a STOCK Xyce whose Linear::Vector / Linear::Matrix objects live in host memory: every
// updateState ships the solution and the store / state vectors to the device, evaluates and assembles there
// (xgpu_load_host: one pass for vectors and matrices), and the load calls add the results into the host objects with
// the "+=" contract of the stock Master.  With device-resident N_LAS objects the same class
// passes device pointers to xgpu_update_state / xgpu_load_vectors / xgpu_load_matrices instead and the copies vanish.
//
// Compiled against the reference headers (it is reference-side code): oracle/Makefile builds it into the oracle
// library, where tests/test_gpu_adaptor.py drives it and the stock Master through the same Device virtuals in one process.
// Lead currents (loadLeadCurrent) are forwarded (xgpu_b4_lead_set / xgpu_lead_load_host).
// Not supported (keep such instances on the stock Master): trnqsMod = 1, IC= rows.
#ifndef Xyce_N_DEV_GpuMaster_B4_h
#define Xyce_N_DEV_GpuMaster_B4_h

#include <map>
#include <string>
#include <vector>

#include <N_DEV_DeviceOptions.h>
#include <N_DEV_ExternData.h>
#include <N_DEV_MOSFET_B4.h>
#include <N_DEV_SolverState.h>
#include <N_LAS_Matrix.h>

#include "../include/xyce_b200.h"
#include "../xyce_b200/csrc/bsim4_fields.def"      // field lists of the records (X-macros named like the reference members)

namespace Xyce {
namespace Device {
namespace MOSFET_B4 {

class GpuMaster : public Master {
 public:
  GpuMaster(const Configuration &configuration, const FactoryBlock &factory_block, const SolverState &ss1, const DeviceOptions &do1)
      : Master(configuration, factory_block, ss1, do1) {}
  ~GpuMaster() { if (ctx_) xgpu_destroy(ctx_); }

  const std::string &lastError() const { return err_; }

  // Once per netlist, after Topology has handed out all LIDs (Instance::registerLIDs / registerStateLIDs /
  // registerStoreLIDs / registerJacLIDs, setupPointers; N_CIR_Xyce.C:667): extracts what Xyce's own host code computed
  // (Model::processParams, SizeDependParam bins, Instance::processParams / updateTemperature) into the records of the
  // C ABI and builds the device-side maps.  ground_lid: the LID that stands for ground in this build (-1 in Xyce).
  bool attach(int cuda_device, int n_unknowns, int ground_lid, int n_state, int n_store) {
    if (xgpu_create(cuda_device, &ctx_) != 0) { err_ = "xgpu_create failed (a CUDA device is required)"; return false; }
    n_ = n_unknowns; nState_ = n_state; nStore_ = n_store;
    std::vector<double> md, sd, id;
    std::vector<int32_t> mi, ii, midx, sidx, lids, sto0, sta0;
    std::map<const Model *, int> mmap;
    std::map<const SizeDependParam *, int> smap;
    // instances sorted by (model, bin): runs of equal records take the uniform-record kernel
    std::vector<Instance *> order(getInstanceBegin(), getInstanceEnd());
    std::stable_sort(order.begin(), order.end(), [](const Instance *a, const Instance *b) {
      return std::make_pair((const void *)&a->model_, (const void *)a->paramPtr) < std::make_pair((const void *)&b->model_, (const void *)b->paramPtr); });
    order_ = order;
    for (Instance *ip : order) {
      Instance &in = *ip;
      Model &mo = in.model_;
      const SizeDependParam &sp = *in.paramPtr;
      if (in.trnqsMod) { err_ = "trnqsMod = 1 is not supported by the GPU master"; return false; }
      if (!mmap.count(&mo)) {
        const int k = (int)mmap.size(); mmap[&mo] = k;
#define XB_PUT(n) md.push_back(mo.n);
        XB_B4_MODEL_D(XB_PUT)
#undef XB_PUT
#define XB_PUT(n) mi.push_back((int32_t)mo.n);
        XB_B4_MODEL_I(XB_PUT)
#undef XB_PUT
      }
      if (!smap.count(&sp)) {
        const int k = (int)smap.size(); smap[&sp] = k;
#define XB_PUT(n) sd.push_back(sp.n);
        XB_B4_SIZE_D(XB_PUT)
#undef XB_PUT
      }
      const size_t i0 = id.size();
#define XB_PUT(n) id.push_back(in.n);
      XB_B4_INST_D(XB_PUT)
#undef XB_PUT
      // toxp / coxp are instance members from 4.8.2 on (N_DEV_MOSFET_B4.h:584-585); the 4.7.0 / 4.6.1 evaluators read the
      // model's (N_DEV_MOSFET_B4p70.C:4527, :4537)
      if (mo.versionDouble < 4.8) {
        size_t k = i0;
#define XB_PUT(n) if (std::string(#n) == "toxp") id[k] = mo.toxp; if (std::string(#n) == "coxp") id[k] = mo.coxp; ++k;
        XB_B4_INST_D(XB_PUT)
#undef XB_PUT
      }
#define XB_PUT(n) ii.push_back((int32_t)in.n);
      XB_B4_INST_I(XB_PUT)
#undef XB_PUT
      midx.push_back(mmap[&mo]); sidx.push_back(smap[&sp]);
      const int l[12] = {in.li_Drain, in.li_GateExt, in.li_Source, in.li_Body, in.li_DrainPrime, in.li_SourcePrime,
                         in.li_GatePrime, in.li_GateMid, in.li_BodyPrime, in.li_SourceBody, in.li_DrainBody, ground_lid};
      for (int t = 0; t < 12; ++t) lids.push_back(l[t] == ground_lid ? -1 : l[t]);
      sto0.push_back(in.li_store_vbd);       // the 22 store slots and the state slots are consecutive LIDs (registerStoreLIDs :6445-6487)
      sta0.push_back(in.li_state_qb);
      const int b0 = in.loadLeadCurrent ? in.li_branch_dev_id : -1;      // id, ig, is, ib (registerBranchDataLIDs :6482-6500)
      branch0_.push_back(b0);
      if (b0 >= 0) nBranch_ = std::max(nBranch_, b0 + 4);
    }
    if (order.empty()) { err_ = "no BSIM4 instances"; return false; }
    if (!chk(xgpu_b4_models_set(ctx_, (int)mmap.size(), md.data(), mi.data(), (int)smap.size(), sd.data()))) return false;
    if (!chk(xgpu_sizes_set(ctx_, n_state, n_store))) return false;
    if (xgpu_b4_group_add(ctx_, (int)order.size(), id.data(), ii.data(), midx.data(), sidx.data(), lids.data(), sto0.data(), 1,
                          sta0.data(), 1) < 0) { err_ = xgpu_last_error(ctx_); return false; }
    if (nBranch_ > 0 && !chk(xgpu_b4_lead_set(ctx_, 0, branch0_.data()))) return false;
    if (!chk(xgpu_pattern_build(ctx_, n_unknowns))) return false;          // the union of this device type's stamps
    if (!chk(xgpu_finalize(ctx_))) return false;
    nnz_ = xgpu_pattern_nnz(ctx_);
    rowptr_.resize(n_ + 1); colind_.resize(nnz_);
    if (!chk(xgpu_pattern_get(ctx_, rowptr_.data(), colind_.data()))) return false;
    for (std::vector<double> *v : {&f_, &q_, &fl_, &ql_}) v->assign(n_, 0.0);
    dF_.assign(nnz_, 0.0); dQ_.assign(nnz_, 0.0);
    return true;
  }

  // Instance::von (the limiter threshold carried from one evaluation to the next) lives in the GPU context; after a
  // restart the host values are pushed here (values in instance-vector order).
  bool setVon(const double *von_by_instance) {
    std::map<const Instance *, int> pos;
    int k = 0;
    for (InstanceVector::const_iterator it = getInstanceBegin(); it != getInstanceEnd(); ++it) pos[*it] = k++;
    std::vector<double> v(order_.size());
    for (size_t i = 0; i < order_.size(); ++i) v[i] = von_by_instance[pos[order_[i]]];
    return chk(xgpu_b4_von_set(ctx_, 0, v.data()));
  }

  // Device::updateState (Core/N_DEV_Device.h:312): evaluates every instance at solVec and publishes the store / state
  // vectors exactly as Master::updateState does (N_DEV_MOSFET_B4.C:10540-10670).
  bool updateState(double *solVec, double *staVec, double *stoVec) override {
    if (!ctx_) return false;
    const SolverState &s = getSolverState();
    const DeviceOptions &o = getDeviceOptions();
    // initial junction voltages taken from an operating-point file (flagSolVectorPtr) are not restated: fail loudly
    if (s.inputOPFlag) { err_ = "inputOPFlag (operating point read from a file) is not supported by the GPU master"; return false; }
    xgpu_solver_state ss;
    ss.dcopFlag = s.dcopFlag; ss.tranopFlag = s.tranopFlag; ss.acopFlag = s.acopFlag; ss.transientFlag = s.transientFlag;
    ss.dcsweepFlag = s.dcsweepFlag; ss.initJctFlag = s.initJctFlag_; ss.initFixFlag = s.initFixFlag; ss.initTranFlag = s.initTranFlag_;
    ss.newtonIter = s.newtonIter; ss.locaEnabledFlag = s.locaEnabledFlag; ss.artParameterFlag = s.artParameterFlag_;
    ss.voltageLimiterFlag = o.voltageLimiterFlag;
    ss.gmin = o.gmin; ss.gainScale = s.gainScale_; ss.nltermScale = s.nltermScale_; ss.vgstConst = o.vgstConst; ss.vdsScaleMin = o.vdsScaleMin;
    ss.sizeScale = s.sizeScale_; ss.currTimeStep = s.currTimeStep_;
    ss.lastTimeStep = s.lastTimeStep_; ss.beginIntegrationFlag = s.beginIntegrationFlag_;
    const ExternData &e = extData_();
    // host-resident DataStore: the time integrator rotates curr / next on the host, so both travel every time
    if (!chk(xgpu_state_set(ctx_, 0, e.nextStoVectorRawPtr)) || !chk(xgpu_state_set(ctx_, 1, e.currStoVectorRawPtr))) return false;
    if (nState_ > 0 && (!chk(xgpu_state_set(ctx_, 2, e.nextStaVectorRawPtr)) || !chk(xgpu_state_set(ctx_, 3, e.currStaVectorRawPtr)))) return false;
    if (!chk(xgpu_load_host(ctx_, solVec, &ss, f_.data(), q_.data(), fl_.data(), ql_.data(), dF_.data(), dQ_.data()))) return false;
    if (!chk(xgpu_state_get(ctx_, 0, e.nextStoVectorRawPtr))) return false;
    if (nState_ > 0) {
      if (!chk(xgpu_state_get(ctx_, 2, e.nextStaVectorRawPtr))) return false;
      if (!s.dcopFlag && s.initTranFlag_ && s.newtonIter == 0 && !chk(xgpu_state_get(ctx_, 3, e.currStaVectorRawPtr))) return false;
    }
    (void)staVec; (void)stoVec;       // the stock Master ignores them too and goes through ExternData (N_DEV_MOSFET_B4.C:10552)
    return true;
  }

  // Device::loadDAEVectors (N_DEV_Device.h:378): "+=" into F, Q and the voltage-limiter vectors
  bool loadDAEVectors(double *solVec, double *fVec, double *qVec, double *bVec, double *leadF, double *leadQ, double *junctionV) override {
    (void)solVec; (void)bVec;
    if (nBranch_ > 0 && leadF && leadQ && junctionV && !chk(xgpu_lead_load_host(ctx_, nBranch_, leadF, leadQ, junctionV))) return false;
    const ExternData &e = extData_();
    for (int i = 0; i < n_; ++i) { fVec[i] += f_[i]; qVec[i] += q_[i]; }
    if (getDeviceOptions().voltageLimiterFlag) {
      double *dFdxdVp = e.dFdxdVpVectorRawPtr, *dQdxdVp = e.dQdxdVpVectorRawPtr;
      for (int i = 0; i < n_; ++i) { dFdxdVp[i] += fl_[i]; dQdxdVp[i] += ql_[i]; }
    }
    return true;
  }

  // Device::loadDAEMatrices (N_DEV_Device.h:427): "+=" into dFdx, dQdx through the matrices' own element addressing
  bool loadDAEMatrices(Linear::Matrix &dFdx, Linear::Matrix &dQdx) override {
    if (pF_.empty()) {      // entry pointers, looked up once (Matrix::returnRawEntryPointer, N_LAS_Matrix.h)
      pF_.resize(nnz_); pQ_.resize(nnz_);
      for (int r = 0; r < n_; ++r)
        for (int k = rowptr_[r]; k < rowptr_[r + 1]; ++k) {
          pF_[k] = dFdx.returnRawEntryPointer(r, colind_[k]);
          pQ_[k] = dQdx.returnRawEntryPointer(r, colind_[k]);
        }
    }
    for (int k = 0; k < nnz_; ++k) { *pF_[k] += dF_[k]; *pQ_[k] += dQ_[k]; }
    return true;
  }

  // Device::isConverged (N_DEV_Device.h:531): AND of Instance::isConverged() = !limitedFlag (N_DEV_MOSFET_B4.h:2328-2331)
  bool isConverged() const override {
    int c = 1;
    return xgpu_all_converged(ctx_, &c) == 0 && c != 0;
  }

 private:
  const ExternData &extData_() const { return (*getInstanceBegin())->extData; }
  bool chk(int rc) { if (rc != 0) { err_ = ctx_ ? xgpu_last_error(ctx_) : "no context"; return false; } return true; }
  xgpu_ctx *ctx_ = nullptr;
  int n_ = 0, nnz_ = 0, nState_ = 0, nStore_ = 0, nBranch_ = 0;
  std::vector<int32_t> branch0_;      // first branch-data LID of every instance (sorted order), -1 = no lead currents
  std::vector<int32_t> rowptr_, colind_;
  std::vector<double> f_, q_, fl_, ql_, dF_, dQ_;
  std::vector<double *> pF_, pQ_;
  std::vector<Instance *> order_;      // upload order (sorted by model, bin)
  std::string err_;
};

}  // namespace MOSFET_B4
}  // namespace Device
}  // namespace Xyce
#endif
