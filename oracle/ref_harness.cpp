// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// C-ABI harness around the *reference's own* device-model objects.  The BSIM4
// translation units are compiled where they lie under /root/reference (see
// oracle/Makefile) and linked with oracle/shim/xyce_shim.cpp; this file drives them
// exactly as Xyce's loader does:
//   Config<Traits>::addConfiguration -> Traits::factory (Master) -> addModel/addInstance
//   registerLIDs / registerStateLIDs / registerStoreLIDs / registerJacLIDs / setupPointers
//     (Topology: N_TOP_CktGraphBasic.C:374-416, :588-637; N_TOP_Indexor.C:149-214)
//   Master::updateState -> loadDAEVectors -> loadDAEMatrices
//     (DeviceMgr: Core/N_DEV_DeviceMgr.C:3857-3936, :4156-4284, :3980-4115)
// on CSR-backed Linear::Matrix objects with the N_LAS_EpetraMatrix.C:658-664 addressing
// rule (row base + offset, ground -> scratch).  It also exports the instance / bin /
// model constants that Xyce's host code (processParams / updateTemperature) computed,
// which is what a real adaptor would upload to the GPU, and the reference's cached
// intermediates for name-by-name comparison.
#include <Xyce_config.h>
#include <algorithm>
#include <cstring>
#include <iostream>
#include <map>
#include <set>
#include <string>
#include <vector>

// compiled with -fno-access-control (oracle/Makefile): the harness reads private members
#include <N_DEV_MOSFET_B4.h>
#include <N_DEV_MOSFET1.h>
#include <N_DEV_Diode.h>
#include <N_DEV_BJT.h>
#include <N_DEV_ADMSmvs_2_0_0_etsoi.h>
// models that went through the ADMS translator (xyce_b200/adms/translate.py): the reference's generated classes
// + the fillers that copy their members into the flat records
#if __has_include("gen_adms/oracle_registry.h")
#include "gen_adms/oracle_registry.h"
#define XB_HAVE_ADMS_ORACLE 1
#endif
#include <N_DEV_Configuration.h>
#include <N_DEV_DeviceBlock.h>
#include <N_DEV_DeviceMaster.h>
#include <N_DEV_DeviceOptions.h>
#include <N_DEV_ExternData.h>
#include <N_DEV_MatrixLoadData.h>
#include <N_LAS_Vector.h>
#include <N_DEV_SolverState.h>
#include <N_LAS_Matrix.h>
#include <N_UTL_Math.h>
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

#include "../adaptor/N_DEV_GpuMaster_Simple.h"  // the same for Diode / MOSFET level 1 / BJT and for translated ADMS models
#include "../adaptor/N_DEV_GpuMaster_B4.h"      // the Xyce-side adaptor (product code for the reference tree), driven by tests/test_gpu_adaptor.py
#include "../xyce_b200/csrc/bsim4_fields.def"
#include "b4_mid_members.def"
#include "../xyce_b200/csrc/tran_driver.h"   // control flow only (Newton / OneStep restatement)

using namespace Xyce;
using namespace Xyce::Device;

namespace {

// CSR matrix with the reference's element addressing semantics.
class CsrMatrix : public Linear::Matrix {
 public:
  std::vector<int> rowptr, colind;
  std::vector<double> vals;
  void setGround(int lid) { groundLID_ = lid; }
  double *operator()(int row, int off) override {
    if (row != groundLID_ && row >= 0 && off >= 0) return &vals[rowptr[row] + off];
    return &groundNode_;
  }
  const double *operator()(int row, int off) const override {
    if (row != groundLID_ && row >= 0 && off >= 0) return &vals[rowptr[row] + off];
    return &groundNode_;
  }
  double *returnRawEntryPointer(int r, int c) override {
    if (r == groundLID_ || r < 0 || c < 0) return &groundNode_;
    for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) if (colind[k] == c) return &vals[k];
    return &groundNode_;
  }
  void put(double s) override { std::fill(vals.begin(), vals.end(), s); }
  // unused parts of the interface
  int setUseTranspose(bool) override { return 0; }
  bool useTranspose() const override { return false; }
  bool putRow(int, int, const double *, const int *) override { return false; }
  void add(const Linear::Matrix &) override {}
  void matvec(bool, const Linear::MultiVector &, Linear::MultiVector &) const override {}
  void linearCombo(const double, const Linear::Matrix &, const double, const Linear::Matrix &) override {}
  void getDiagonal(Linear::Vector &) const override {}
  bool replaceDiagonal(const Linear::Vector &) override { return false; }
  int getRowLength(int r) const override { return rowptr[r + 1] - rowptr[r]; }
  int getLocalRowLength(int r) const override { return rowptr[r + 1] - rowptr[r]; }
  int getLocalRowView(int, int &, double *&, int *&) const override { return -1; }
  void getRowCopy(int, int, int &, double *, int *) const override {}
  void getLocalRowCopy(int, int, int &, double *, int *) const override {}
  int getNumRows() const override { return (int)rowptr.size() - 1; }
  int getLocalNumRows() const override { return (int)rowptr.size() - 1; }
  bool addIntoLocalRow(int, int, const double *, const int *) override { return false; }
  bool putLocalRow(int, int, const double *, const int *) override { return false; }
  void writeToFile(const char *, bool, bool) const override {}
  const Parallel::ParMap *getColMap(const Parallel::Communicator &) const override { return 0; }
  const Linear::Graph *getGraph() const override { return 0; }
  void scale(double) override {}
  void print(std::ostream &) const override {}
};

// View of a std::vector<double> with the Linear::Vector element-access contract (operator[] on the
// overlapped storage, N_LAS_EpetraVector.h:183-198); MOSFET level 1 and the BJT read their solution and
// store/state history through these objects rather than through the raw pointers.
class RawVector : public Linear::Vector {
 public:
  std::vector<double> *v = nullptr;
  double &operator[](int i) override { return (*v)[i]; }
  const double &operator[](int i) const override { return (*v)[i]; }
  Linear::Vector &operator=(const Linear::Vector &) override { return *this; }
  Linear::MultiVector &operator=(const Linear::MultiVector &) override { return *this; }
  Linear::Vector *cloneVector() const override { return 0; }
  Linear::Vector *cloneCopyVector() const override { return 0; }
  double dotProduct(const Linear::Vector &) const override { return 0; }
  Linear::MultiVector *clone() const override { return 0; }
  Linear::MultiVector *cloneCopy() const override { return 0; }
  void dotProduct(const Linear::MultiVector &, std::vector<double> &) const override {}
  void scale(const double) override {}
  void multiply(const Linear::MultiVector &) override {}
  void update(double, const Linear::MultiVector &, double) override {}
  void update(double, const Linear::MultiVector &, double, const Linear::MultiVector &, double) override {}
  int lpNorm(const int, double *) const override { return 0; }
  int infNorm(double *, int *) const override { return 0; }
  int wRMSNorm(const Linear::MultiVector &, double *) const override { return 0; }
  int wMaxNorm(const Linear::MultiVector &, double *, int *) const override { return 0; }
  void random() override {}
  void putScalar(const double a) override { std::fill(v->begin(), v->end(), a); }
  void addScalar(const double) override {}
  void absValue(const Linear::MultiVector &) override {}
  void reciprocal(const Linear::MultiVector &) override {}
  double *operator()(int r, int) override { return &(*v)[r]; }
  const double *operator()(int r, int) const override { return &(*v)[r]; }
  const Linear::Vector *getVectorView(int) const override { return this; }
  Linear::Vector *getNonConstVectorView(int) override { return this; }
  const Linear::Vector *getVectorViewAssembled(int) const override { return this; }
  Linear::Vector *getNonConstVectorViewAssembled(int) override { return this; }
  int globalLength() const override { return (int)v->size(); }
  int localLength() const override { return (int)v->size(); }
  int numVectors() const override { return 1; }
  int externVectorSize() const override { return 0; }
  bool vectorImport(const Linear::MultiVector *, Linear::Importer *) override { return false; }
  bool importOverlap() override { return false; }
  void writeToFile(const char *, bool, bool) const override {}
  const double &getElementByGlobalIndex(const int &i, const int &) const override { return (*v)[i]; }
  bool setElementByGlobalIndex(const int &, const double &, const int &) override { return false; }
  bool sumElementByGlobalIndex(const int &, const double &, const int &) override { return false; }
  void clearExternVectorMap() override {}
  void addElementToExternVectorMap(const int &, const double &) override {}
  void print(std::ostream &) const override {}
  const Parallel::ParMap *pmap() const override { return 0; }
  const Parallel::ParMap *omap() const override { return 0; }
  const Parallel::Communicator *pdsComm() const override { return 0; }
  void fillComplete() override {}
};

struct InstRec {
  DeviceInstance *inst;
  int dev;                   // index into Ctx::masters
  std::vector<int> ext;      // external node ids (-1 = ground)
  std::vector<int> lids;     // ext + int LIDs
  int sta0, sto0;
  int br0 = -1;              // first branch-data LID (lead currents), -1 when not enabled
};

#ifdef XB_HAVE_ADMS_ORACLE
// record fillers of the translated ADMS models for GpuSimpleMaster: the generated adms_fill_<model>() + the type id the
// library's registry gives that model
#define XB_ORACLE_GPUFILL(nm_, ns_) \
  struct GpuAdmsFill_##nm_ { \
    typedef ns_::Instance Inst; \
    static int type() { static const int t = gpu_adms_type_by_name(#nm_); return t; } \
    static int fill(Inst &in, std::vector<double> &rec, std::vector<int32_t> &lids, int &flags) { \
      double r[4096]; int l[64]; \
      const int k = adms_fill_##nm_(in, r, l), nl = adms_nlids_##nm_; \
      rec.insert(rec.end(), r, r + k); lids.insert(lids.end(), l, l + nl); flags = 0; \
      return nl; \
    } \
    static int branch0(Inst &) { return -1; } \
    static int branches() { return 0; } \
  };
XB_ADMS_ORACLE_LIST(XB_ORACLE_GPUFILL)
#undef XB_ORACLE_GPUFILL
#endif

struct Ctx {
  DeviceOptions devOptions;
  SolverState solState;
  ExternData extData;
  RawVector vSol, vCurrSta, vNextSta, vCurrSto, vNextSto, vLastSto, vF, vQ;
  MatrixLoadData mlData;
  FactoryBlock *fb = 0;
  // one Master per device type, in creation order (= DeviceMgr::devicePtrVec_ order)
  std::vector<Xyce::Device::Device *> masters;
  std::vector<std::string> masterType;
  std::vector<Configuration *> masterCfg;
  int master(const std::string &t) {
    for (size_t i = 0; i < masterType.size(); ++i) if (masterType[i] == t) return (int)i;
    Configuration *cfg = 0; Xyce::Device::Device *m = 0;
    if (t == "b4") {
      auto *c = &Config<MOSFET_B4::Traits>::addConfiguration(); cfg = c;
      // the factory hook (N_DEV_MOSFET_B4.C:11688-11692): stock Master, or the GPU adaptor with the same constructor arguments
      if (gpuB4) m = new MOSFET_B4::GpuMaster(*c, *fb, fb->solverState_, fb->deviceOptions_);
      else m = MOSFET_B4::Traits::factory(*c, *fb);
    }
    else if (t == "m1") { auto *c = &Config<MOSFET1::Traits>::addConfiguration(); cfg = c;
      m = gpuAll ? static_cast<Xyce::Device::Device *>(new GpuMos1Master(*c, *fb, fb->solverState_, fb->deviceOptions_)) : MOSFET1::Traits::factory(*c, *fb); }
    else if (t == "d") { auto *c = &Config<Diode::Traits>::addConfiguration(); cfg = c;
      m = gpuAll ? static_cast<Xyce::Device::Device *>(new GpuDiodeMaster(*c, *fb, fb->solverState_, fb->deviceOptions_)) : Diode::Traits::factory(*c, *fb); }
    else if (t == "q") { auto *c = &Config<BJT::Traits>::addConfiguration(); cfg = c;
      m = gpuAll ? static_cast<Xyce::Device::Device *>(new GpuBjtMaster(*c, *fb, fb->solverState_, fb->deviceOptions_)) : BJT::Traits::factory(*c, *fb); }
    else if (t == "mvs") { auto *c = &Config<ADMSmvs_2_0_0_etsoi::Traits>::addConfiguration(); cfg = c; m = ADMSmvs_2_0_0_etsoi::Traits::factory(*c, *fb); }
#ifdef XB_HAVE_ADMS_ORACLE
#define XB_ORACLE_MASTER(nm_, ns_) else if (t == "adms:" #nm_) { auto *c = &Config<ns_::Traits>::addConfiguration(); cfg = c; \
      m = gpuAll ? static_cast<Xyce::Device::Device *>(new GpuSimpleMaster<DeviceMaster<ns_::Traits>, GpuAdmsFill_##nm_>(*c, *fb, fb->solverState_, fb->deviceOptions_)) \
                 : ns_::Traits::factory(*c, *fb); }
    XB_ADMS_ORACLE_LIST(XB_ORACLE_MASTER)
#undef XB_ORACLE_MASTER
#endif
    else return -1;
    masters.push_back(m); masterType.push_back(t); masterCfg.push_back(cfg);
    return (int)masters.size() - 1;
  }
  bool updateAll() {
    bool ok = true;
    for (auto *m : masters) ok = m->updateState(sol.data(), nextSta.data(), nextSto.data()) && ok;
    for (auto *m : masters) ok = m->updateSecondaryState(staDeriv.data(), nextSto.data()) && ok;
    return ok;
  }
  bool loadVectorsAll() {
    bool ok = true;
    double *lf = leadF.empty() ? 0 : leadF.data(), *lq = leadQ.empty() ? 0 : leadQ.data(), *jv = junctionV.empty() ? 0 : junctionV.data();
    for (auto *m : masters) ok = m->loadDAEVectors(sol.data(), f.data(), q.data(), b.data(), lf, lq, jv) && ok;
    return ok;
  }
  bool loadMatricesAll() {
    bool ok = true;
    for (auto *m : masters) ok = m->loadDAEMatrices(dFdx, dQdx) && ok;
    return ok;
  }
  bool gpuAll = false;       // Diode / MOSFET1 / BJT / translated ADMS instances go to GpuSimpleMaster (adaptor/N_DEV_GpuMaster_Simple.h)
  bool gpuB4 = false;        // BSIM4 instances go to MOSFET_B4::GpuMaster (adaptor/N_DEV_GpuMaster_B4.h) instead of the stock Master
  std::vector<double> staDeriv, leadF, leadQ, junctionV;
  bool lead = false;         // DeviceInstance::enableLeadCurrentCalc on every instance (what .PRINT I(...) / P(...) triggers)
  std::vector<InstRec> insts;
  int nExt = 0, n = 0, nSta = 0, nSto = 0;
  CsrMatrix dFdx, dQdx;
  std::vector<double> f, q, b, fl, ql, nextSta, currSta, nextSto, currSto, lastSto, sol;
  bool finalized = false;
  std::vector<std::pair<int, int>> extra_pattern;   // linear-device entries
};

alignas(64) char g_fake_mgr[4096];
alignas(64) char g_fake_cmd[4096];

std::vector<Param> make_params(int n, const char **keys, const double *vals) {
  std::vector<Param> v;
  for (int i = 0; i < n; ++i) {
    Param p(std::string(keys[i]), vals[i], true);
    v.push_back(p);
  }
  return v;
}

}  // namespace

static void xref_segv(int) { void *bt[64]; int n = backtrace(bt, 64); backtrace_symbols_fd(bt, n, 2); _exit(139); }

extern "C" {

void *xref_new() {
  if (getenv("XREF_DEBUG")) signal(SIGSEGV, xref_segv);
  Ctx *c = new Ctx;
  c->fb = new FactoryBlock(*reinterpret_cast<DeviceMgr *>(g_fake_mgr), c->devOptions, c->solState, c->mlData,
                           c->extData, *reinterpret_cast<IO::CmdParse *>(g_fake_cmd));
  return c;
}

int xref_set_num_external_nodes(void *h, int n) { ((Ctx *)h)->nExt = n; return 0; }

// devtype: "b4" BSIM4 (level 54), "m1" MOSFET level 1, "d" diode, "q" Gummel-Poon BJT, "mvs" ADMS-generated MVS 2.0.0 ETSOI.
// mtype: the .model type string in upper case (NMOS, PMOS, D, NPN, PNP).
int xref_add_model(void *h, const char *devtype, const char *name, const char *mtype, int level, int np,
                   const char **keys, const double *vals) {
  Ctx *c = (Ctx *)h;
  const int d = c->master(devtype);
  if (d < 0) return 2;
  ModelBlock mb(name, mtype, level);
  mb.params = make_params(np, keys, vals);
  return c->masters[d]->addModel(mb, *c->fb) ? 0 : 1;
}

int xref_add_instance(void *h, const char *devtype, const char *name, const char *model, int nnodes, const int *nodes,
                      int np, const char **keys, const double *vals) {
  Ctx *c = (Ctx *)h;
  const int d = c->master(devtype);
  if (d < 0) return 2;
  InstanceBlock ib{std::string(name)};
  ib.setModelName(ModelName(model));
  ib.params = make_params(np, keys, vals);
  ib.iNumNodes = nnodes;
  ib.numExtVars = nnodes;
  ib.modelFlag = true;
  DeviceInstance *di = c->masters[d]->addInstance(ib, *c->fb);
  if (!di) return 1;
  InstRec r;
  r.inst = di;
  r.dev = d;
  r.ext.assign(nodes, nodes + nnodes);
  c->insts.push_back(r);
  return 0;
}

int xref_b4_add_model(void *h, const char *name, const char *type, int np, const char **keys, const double *vals) {
  return xref_add_model(h, "b4", name, type, 54, np, keys, vals);
}
int xref_b4_add_instance(void *h, const char *name, const char *model, const int *nodes4, int np,
                         const char **keys, const double *vals) {
  return xref_add_instance(h, "b4", name, model, 4, nodes4, np, keys, vals);
}

// Assign LIDs, build the CSR pattern from every instance's jacobianStamp(), register
// everything with the instances.  Returns number of unknowns.
int xref_finalize(void *h) {
  Ctx *c = (Ctx *)h;
  int next = c->nExt;
  for (auto &r : c->insts) {
    r.lids = r.ext;
    for (int k = 0; k < r.inst->getNumIntVars(); ++k) r.lids.push_back(next++);
  }
  c->n = next;
  const int ground = c->n;
  std::vector<std::set<int>> rows(c->n);
  for (auto &r : c->insts) {
    const std::vector<std::vector<int>> &st = r.inst->jacobianStamp();
    for (size_t i = 0; i < st.size(); ++i) {
      int gr = r.lids[i];
      if (gr < 0) continue;
      for (int cj : st[i]) { int gc = r.lids[cj]; if (gc >= 0) rows[gr].insert(gc); }
    }
  }
  for (auto &e : c->extra_pattern) rows[e.first].insert(e.second);
  for (CsrMatrix *M : {&c->dFdx, &c->dQdx}) {
    M->rowptr.assign(1, 0);
    M->colind.clear();
    for (int i = 0; i < c->n; ++i) {
      for (int cc : rows[i]) M->colind.push_back(cc);
      M->rowptr.push_back((int)M->colind.size());
    }
    M->vals.assign(M->colind.size(), 0.0);
    M->setGround(ground);
  }
  int sta = 0, sto = 0, nbr = 0;
  for (auto &r : c->insts) {
    std::vector<int> ext, in;
    const int ne = r.inst->getNumExtVars();
    for (int k = 0; k < ne; ++k) ext.push_back(r.lids[k] < 0 ? ground : r.lids[k]);
    for (size_t k = ne; k < r.lids.size(); ++k) in.push_back(r.lids[k]);
    r.inst->registerLIDs(in, ext);
    std::vector<int> sv, tv;
    r.sta0 = sta; r.sto0 = sto;
    for (int k = 0; k < r.inst->getNumStateVars(); ++k) sv.push_back(sta++);
    for (int k = 0; k < r.inst->getNumStoreVars(); ++k) tv.push_back(sto++);
    r.inst->registerStateLIDs(sv);
    r.inst->registerStoreLIDs(tv);
    if (c->lead) {
      r.inst->enableLeadCurrentCalc();
      std::vector<int> bv;
      r.br0 = nbr;
      for (int k = 0; k < r.inst->getNumBranchDataVars(); ++k) bv.push_back(nbr++);
      r.inst->registerBranchDataLIDs(bv);
    }
    const std::vector<std::vector<int>> &st = r.inst->jacobianStamp();
    std::vector<std::vector<int>> jl(st.size());
    for (size_t i = 0; i < st.size(); ++i) {
      jl[i].assign(st[i].size(), -1);
      int gr = r.lids[i];
      if (gr < 0) continue;
      for (size_t j = 0; j < st[i].size(); ++j) {
        int gc = r.lids[st[i][j]];
        if (gc < 0) continue;
        const int *b = &c->dFdx.colind[c->dFdx.rowptr[gr]], *e = &c->dFdx.colind[c->dFdx.rowptr[gr + 1]];
        jl[i][j] = (int)(std::lower_bound(b, e, gc) - b);
      }
    }
    r.inst->registerJacLIDs(jl);
  }
  c->nSta = sta; c->nSto = sto;
  if (c->lead) {
    c->leadF.assign(nbr + 1, 0.0); c->leadQ = c->junctionV = c->leadF;
    c->extData.nextLeadCurrFCompRawPtr = c->leadF.data(); c->extData.nextLeadCurrQCompRawPtr = c->leadQ.data();
    c->extData.nextJunctionVCompRawPtr = c->junctionV.data();
  }
  c->f.assign(c->n + 1, 0); c->q = c->b = c->fl = c->ql = c->f;
  c->sol.assign(c->n + 1, 0);
  c->nextSta.assign(sta + 1, 0); c->currSta = c->nextSta; c->staDeriv = c->nextSta;
  c->nextSto.assign(sto + 1, 0); c->currSto = c->nextSto; c->lastSto = c->nextSto;
  ExternData &e = c->extData;
  e.dFdxMatrixPtr = &c->dFdx; e.dQdxMatrixPtr = &c->dQdx;
  e.daeFVectorRawPtr = c->f.data(); e.daeQVectorRawPtr = c->q.data(); e.daeBVectorRawPtr = c->b.data();
  e.dFdxdVpVectorRawPtr = c->fl.data(); e.dQdxdVpVectorRawPtr = c->ql.data();
  e.nextSolVectorRawPtr = e.currSolVectorRawPtr = e.lastSolVectorRawPtr = c->sol.data();
  e.nextStaVectorRawPtr = c->nextSta.data(); e.currStaVectorRawPtr = e.lastStaVectorRawPtr = c->currSta.data();
  e.nextStaDerivVectorRawPtr = c->staDeriv.data();
  e.nextStoVectorRawPtr = c->nextSto.data(); e.currStoVectorRawPtr = c->currSto.data(); e.lastStoVectorRawPtr = c->lastSto.data();
  c->vSol.v = &c->sol; c->vCurrSta.v = &c->currSta; c->vNextSta.v = &c->nextSta; c->vCurrSto.v = &c->currSto; c->vNextSto.v = &c->nextSto; c->vLastSto.v = &c->lastSto;
  e.nextSolVectorPtr = e.currSolVectorPtr = e.lastSolVectorPtr = &c->vSol;
  c->vF.v = &c->f; c->vQ.v = &c->q;
  e.daeFVectorPtr = &c->vF; e.daeQVectorPtr = &c->vQ;      // ADMS-generated devices load through the Linear::Vector objects
  e.currStaVectorPtr = e.lastStaVectorPtr = &c->vCurrSta; e.nextStaVectorPtr = &c->vNextSta;
  e.currStoVectorPtr = &c->vCurrSto; e.lastStoVectorPtr = &c->vLastSto; e.nextStoVectorPtr = &c->vNextSto;
  for (auto &r : c->insts) r.inst->setupPointers();
  c->finalized = true;
  return c->n;
}

// ---- the Xyce-side GPU adaptor in place of the stock BSIM4 Master ----
// xref_use_gpu_master: before the first BSIM4 model / instance is added.  xref_gpu_attach: after xref_finalize
// (all LIDs registered).  From then on updateAll / loadVectorsAll / loadMatricesAll reach the GPU through the same
// Device virtuals the stock Master answers.
// on = 1: BSIM4 only; on = 2: every device type that has an adaptor (BSIM4, diode, MOSFET level 1, BJT, translated ADMS)
void xref_use_gpu_master(void *h, int on) { ((Ctx *)h)->gpuB4 = on != 0; ((Ctx *)h)->gpuAll = on >= 2; }
int xref_gpu_attach(void *h, int cuda_device) {
  Ctx *c = (Ctx *)h;
  int attached = 0;
  for (auto *m : c->masters) {
    if (MOSFET_B4::GpuMaster *g = dynamic_cast<MOSFET_B4::GpuMaster *>(m)) {
      if (!g->attach(cuda_device, c->n, c->n, c->nSta, c->nSto)) { std::cerr << "xref_gpu_attach: " << g->lastError() << std::endl; return 2; }
      ++attached;
    } else if (GpuAttachable *g = dynamic_cast<GpuAttachable *>(m)) {
      if (!g->attach(cuda_device, c->n, c->n, c->nSta, c->nSto)) { std::cerr << "xref_gpu_attach: " << g->lastError() << std::endl; return 2; }
      ++attached;
    }
  }
  return attached > 0 ? 0 : 1;
}
int xref_gpu_set_von(void *h, const double *von) {      // von[i] of the i-th BSIM4 instance (harness order = instance-vector order)
  Ctx *c = (Ctx *)h;
  const int d = c->master("b4");
  MOSFET_B4::GpuMaster *g = d >= 0 ? dynamic_cast<MOSFET_B4::GpuMaster *>(c->masters[d]) : nullptr;
  return (g && g->setVon(von)) ? 0 : 1;
}
int xref_all_converged(void *h) {
  Ctx *c = (Ctx *)h;
  bool ok = true;
  for (auto *m : c->masters) ok = m->isConverged() && ok;      // DeviceMgr::allDevicesConverged (Core/N_DEV_DeviceMgr.C:5603-5640)
  return ok ? 1 : 0;
}

int xref_nnz(void *h) { return (int)((Ctx *)h)->dFdx.colind.size(); }
void xref_pattern(void *h, int *rowptr, int *colind) {
  Ctx *c = (Ctx *)h;
  std::copy(c->dFdx.rowptr.begin(), c->dFdx.rowptr.end(), rowptr);
  std::copy(c->dFdx.colind.begin(), c->dFdx.colind.end(), colind);
}
int xref_num_state(void *h) { return ((Ctx *)h)->nSta; }
int xref_num_store(void *h) { return ((Ctx *)h)->nSto; }

// flags[]: dcop, tranop, acop, transient, dcsweep, initJct, initFix, initTran, newtonIter,
//          locaEnabled, artParameter, voltageLimiter
// dvals[]: gmin, gainScale, nltermScale
void xref_set_flags(void *h, const int *fl, const double *dv) {
  Ctx *c = (Ctx *)h;
  SolverState &s = c->solState;
  s.dcopFlag = fl[0]; s.tranopFlag = fl[1]; s.acopFlag = fl[2]; s.transientFlag = fl[3];
  s.dcsweepFlag = fl[4]; s.initJctFlag_ = fl[5]; s.initFixFlag = fl[6]; s.initTranFlag_ = fl[7];
  s.newtonIter = fl[8]; s.locaEnabledFlag = fl[9]; s.artParameterFlag_ = fl[10];
  c->devOptions.voltageLimiterFlag = fl[11];
  c->devOptions.gmin = dv[0]; s.gainScale_ = dv[1]; s.nltermScale_ = dv[2];
}

// step history read by the BJT excess phase (SolverState::currTimeStep_ / lastTimeStep_ / beginIntegrationFlag_, last store)
void xref_set_step(void *h, double currTimeStep, double lastTimeStep, int beginIntegration) {
  Ctx *c = (Ctx *)h;
  c->solState.currTimeStep_ = currTimeStep; c->solState.lastTimeStep_ = lastTimeStep; c->solState.beginIntegrationFlag_ = beginIntegration != 0;
}
void xref_last_store(void *h, const double *set, double *get) {
  Ctx *c = (Ctx *)h;
  if (set) std::copy(set, set + c->nSto, c->lastSto.begin());
  if (get) std::copy(c->lastSto.begin(), c->lastSto.begin() + c->nSto, get);
}
void xref_set_state(void *h, const double *currSto, const double *nextSto, const double *currSta) {
  Ctx *c = (Ctx *)h;
  if (currSto) std::copy(currSto, currSto + c->nSto, c->currSto.begin());
  if (nextSto) std::copy(nextSto, nextSto + c->nSto, c->nextSto.begin());
  if (currSta) std::copy(currSta, currSta + c->nSta, c->currSta.begin());
}
void xref_get_state(void *h, double *currSto, double *nextSto, double *currSta, double *nextSta) {
  Ctx *c = (Ctx *)h;
  if (currSto) std::copy(c->currSto.begin(), c->currSto.begin() + c->nSto, currSto);
  if (nextSto) std::copy(c->nextSto.begin(), c->nextSto.begin() + c->nSto, nextSto);
  if (currSta) std::copy(c->currSta.begin(), c->currSta.begin() + c->nSta, currSta);
  if (nextSta) std::copy(c->nextSta.begin(), c->nextSta.begin() + c->nSta, nextSta);
}
// carried limiter threshold `von` of every BSIM4 instance
void xref_b4_set_von(void *h, const double *von) {
  Ctx *c = (Ctx *)h;
  for (size_t i = 0; i < c->insts.size(); ++i)
    if (c->masterType[c->insts[i].dev] == "b4") static_cast<MOSFET_B4::Instance *>(c->insts[i].inst)->von = von[i];
}
void xref_b4_get_von(void *h, double *von) {
  Ctx *c = (Ctx *)h;
  for (size_t i = 0; i < c->insts.size(); ++i)
    von[i] = (c->masterType[c->insts[i].dev] == "b4") ? static_cast<MOSFET_B4::Instance *>(c->insts[i].inst)->von : 0.0;
}

// One updateState + loadDAEVectors + loadDAEMatrices pass at solution x[0..n).
// Outputs (each length n, matrices nnz): f, q, dFdxdVp, dQdxdVp, dFdx values, dQdx values.
int xref_load(void *h, const double *x, double *f, double *q, double *fl, double *ql, double *dfdx, double *dqdx) {
  Ctx *c = (Ctx *)h;
  std::copy(x, x + c->n, c->sol.begin());
  c->sol[c->n] = 0.0;
  for (auto *v : {&c->f, &c->q, &c->b, &c->fl, &c->ql}) std::fill(v->begin(), v->end(), 0.0);
  c->dFdx.put(0.0); c->dQdx.put(0.0);
  bool ok = c->updateAll();
  ok = c->loadVectorsAll() && ok;
  ok = c->loadMatricesAll() && ok;
  if (f) std::copy(c->f.begin(), c->f.begin() + c->n, f);
  if (q) std::copy(c->q.begin(), c->q.begin() + c->n, q);
  if (fl) std::copy(c->fl.begin(), c->fl.begin() + c->n, fl);
  if (ql) std::copy(c->ql.begin(), c->ql.begin() + c->n, ql);
  if (dfdx) std::copy(c->dFdx.vals.begin(), c->dFdx.vals.end(), dfdx);
  if (dqdx) std::copy(c->dQdx.vals.begin(), c->dQdx.vals.end(), dqdx);
  return ok ? 0 : 1;
}

// Timing helper for the CPU baseline: `reps` evaluation passes, no output copies.
int xref_load_repeat(void *h, int reps) {
  Ctx *c = (Ctx *)h;
  for (int r = 0; r < reps; ++r) {
    for (auto *v : {&c->f, &c->q, &c->b, &c->fl, &c->ql}) std::fill(v->begin(), v->end(), 0.0);
    c->dFdx.put(0.0); c->dQdx.put(0.0);
    c->updateAll();
    c->loadVectorsAll();
    c->loadMatricesAll();
  }
  return 0;
}
void xref_set_solution(void *h, const double *x) {
  Ctx *c = (Ctx *)h;
  std::copy(x, x + c->n, c->sol.begin());
}

// ---- parameter-record export (what a host adaptor uploads) ----
int xref_b4_counts(int *out) {
#define CNT(n) +1
  out[0] = 0 XB_B4_MODEL_D(CNT); out[1] = 0 XB_B4_MODEL_I(CNT); out[2] = 0 XB_B4_SIZE_D(CNT);
  out[3] = 0 XB_B4_INST_D(CNT); out[4] = 0 XB_B4_INST_I(CNT);
  out[5] = 0 XB_B4_MID_REF_D(CNT); out[6] = 1 XB_B4_MID_REF_I(CNT);
#undef CNT
  return 0;
}
// names, newline separated, in export order: which = 0 model_d,1 model_i,2 size_d,3 inst_d,4 inst_i,5 mid_d,6 mid_i
const char *xref_b4_names(int which) {
#define NM(n) #n "\n"
  switch (which) {
    case 0: return XB_B4_MODEL_D(NM);
    case 1: return XB_B4_MODEL_I(NM);
    case 2: return XB_B4_SIZE_D(NM);
    case 3: return XB_B4_INST_D(NM);
    case 4: return XB_B4_INST_I(NM);
    case 5: return XB_B4_MID_REF_D(NM);
    case 6: return "origFlag\n" XB_B4_MID_REF_I(NM);
  }
#undef NM
  return "";
}
// model/bin identity (pointers as integers) lets the caller de-duplicate records
void xref_b4_export(void *h, int idx, double *model_d, int *model_i, double *size_d, double *inst_d, int *inst_i,
                    long long *model_id, long long *size_id, int *lids12, int *sta0, int *sto0) {
  Ctx *c = (Ctx *)h;
  MOSFET_B4::Instance &in = *static_cast<MOSFET_B4::Instance *>(c->insts[idx].inst);
  MOSFET_B4::Model &mo = in.model_;
  const MOSFET_B4::SizeDependParam &sp = *in.paramPtr;
  int k;
#define PUT(n) model_d[k++] = mo.n;
  k = 0; XB_B4_MODEL_D(PUT)
#undef PUT
#define PUT(n) model_i[k++] = (int)mo.n;
  k = 0; XB_B4_MODEL_I(PUT)
#undef PUT
#define PUT(n) size_d[k++] = sp.n;
  k = 0; XB_B4_SIZE_D(PUT)
#undef PUT
#define PUT(n) inst_d[k++] = in.n;
  k = 0; XB_B4_INST_D(PUT)
#undef PUT
  // toxp / coxp became instance members in 4.8.2 (N_DEV_MOSFET_B4.h:584-585); the 4.7.0 and 4.6.1 evaluators
  // read the model's (N_DEV_MOSFET_B4p70.C:4527, :4537): an adaptor fills the instance record from there
  if (mo.versionDouble < 4.8) {
#define PUT(n) if (!std::strcmp(#n, "toxp")) inst_d[k] = mo.toxp; if (!std::strcmp(#n, "coxp")) inst_d[k] = mo.coxp; ++k;
    k = 0; XB_B4_INST_D(PUT)
#undef PUT
  }
#define PUT(n) inst_i[k++] = (int)in.n;
  k = 0; XB_B4_INST_I(PUT)
#undef PUT
  *model_id = (long long)(size_t)&mo;
  *size_id = (long long)(size_t)&sp;
  const int g = c->n;
  int l[12] = {in.li_Drain, in.li_GateExt, in.li_Source, in.li_Body, in.li_DrainPrime, in.li_SourcePrime,
               in.li_GatePrime, in.li_GateMid, in.li_BodyPrime, in.li_SourceBody, in.li_DrainBody,
               in.trnqsMod ? in.li_Charge : g};
  for (int i = 0; i < 12; ++i) lids12[i] = (l[i] == g) ? -1 : l[i];
  *sta0 = c->insts[idx].sta0;
  *sto0 = c->insts[idx].sto0;
}
// generic: LIDs of all variables (external then internal, -1 = ground), first state / store LID
int xref_inst_info(void *h, int idx, int *lids, int max_lids, int *sta0, int *sto0, int *nsta, int *nsto) {
  Ctx *c = (Ctx *)h;
  const InstRec &r = c->insts[idx];
  const int k = (int)r.lids.size();
  for (int i = 0; i < k && i < max_lids; ++i) lids[i] = r.lids[i];
  *sta0 = r.sta0; *sto0 = r.sto0; *nsta = r.inst->getNumStateVars(); *nsto = r.inst->getNumStoreVars();
  return k;
}

// Diode record in the order of xyce_b200/csrc/diode_eval.h (XB_DIODE_D) + flag word
#include "../xyce_b200/csrc/simple_fields.def"
int xref_diode_export(void *h, int idx, double *rec, int *flags, int *lids3) {
  Ctx *c = (Ctx *)h;
  Diode::Instance &in = *static_cast<Diode::Instance *>(c->insts[idx].inst);
  Diode::Model &mo = in.model_;
  int k = 0;
#define MOD(n) rec[k++] = mo.n;
#define INS(n) rec[k++] = in.n;
  XB_DIODE_FIELDS(MOD, INS)
#undef MOD
#undef INS
  *flags = (mo.BVGiven ? 1 : 0) | (mo.JSWGiven ? 2 : 0) | (mo.NSGiven ? 4 : 0) | (in.InitCondGiven ? 8 : 0) | (in.off ? 16 : 0);
  const int g = c->n;
  const int l[3] = {in.li_Pos, in.li_Neg, in.li_Pri};
  for (int i = 0; i < 3; ++i) lids3[i] = (l[i] == g) ? -1 : l[i];
  return k;
}

// ADMS-generated MVS 2.0.0 ETSOI: model-card record in XB_MVS_FIELDS order + the 7 unknown LIDs (d g s di si sf branch)
int xref_mvs_export(void *h, int idx, double *rec, int *lids7) {
  Ctx *c = (Ctx *)h;
  ADMSmvs_2_0_0_etsoi::Instance &in = *static_cast<ADMSmvs_2_0_0_etsoi::Instance *>(c->insts[idx].inst);
  ADMSmvs_2_0_0_etsoi::Model &mo = in.model_;
  int k = 0;
#define X(n) rec[k++] = mo.n;
  XB_MVS_FIELDS(X)
#undef X
  const int g = c->n;
  const int l[7] = {in.li_d, in.li_g, in.li_s, in.li_di, in.li_si, in.li_sf, in.li_BRA_sf_GND};
  for (int i = 0; i < 7; ++i) lids7[i] = (l[i] == g) ? -1 : l[i];
  return k;
}

// Any model of the ADMS translator's registry: record in the evaluator's field order + the unknowns' LIDs (-1 = ground)
int xref_adms_export(void *h, int idx, const char *name, double *rec, int *lids, int *n_lids) {
  Ctx *c = (Ctx *)h;
  const std::string nm(name);
  int k = -1, nl = 0;
#ifdef XB_HAVE_ADMS_ORACLE
#define XB_ORACLE_FILL(nm_, ns_) if (nm == #nm_) { ns_::Instance &in = *static_cast<ns_::Instance *>(c->insts[idx].inst); \
    k = adms_fill_##nm_(in, rec, lids); nl = adms_nlids_##nm_; }
  XB_ADMS_ORACLE_LIST(XB_ORACLE_FILL)
#undef XB_ORACLE_FILL
#endif
  for (int i = 0; i < nl; ++i) if (lids[i] == c->n) lids[i] = -1;
  *n_lids = nl;
  return k;
}

// lead currents (loadLeadCurrent): call before xref_finalize; vectors are indexed by branch-data LID
void xref_enable_lead_currents(void *h) { ((Ctx *)h)->lead = true; }
int xref_num_branch_data(void *h) { return (int)((Ctx *)h)->leadF.size() - 1; }
int xref_inst_branch0(void *h, int idx) { return ((Ctx *)h)->insts[idx].br0; }
void xref_get_lead(void *h, double *leadF, double *leadQ, double *junctionV) {
  Ctx *c = (Ctx *)h;
  const size_t n = c->leadF.size() - 1;
  std::copy(c->leadF.begin(), c->leadF.begin() + n, leadF);
  std::copy(c->leadQ.begin(), c->leadQ.begin() + n, leadQ);
  std::copy(c->junctionV.begin(), c->junctionV.begin() + n, junctionV);
}

int xref_inst_converged(void *h, int idx) { return ((Ctx *)h)->insts[idx].inst->isConverged() ? 1 : 0; }

// MOSFET level 1 / BJT records in the order of XB_MOS1_FIELDS / XB_BJT_FIELDS + flag word + node LIDs
int xref_mos1_export(void *h, int idx, double *rec, int *flags, int *lids6) {
  Ctx *c = (Ctx *)h;
  MOSFET1::Instance &in = *static_cast<MOSFET1::Instance *>(c->insts[idx].inst);
  MOSFET1::Model &mo = in.getModel();
  int k = 0;
#define MOD(n) rec[k++] = mo.n;
#define INS(n) rec[k++] = in.n;
  XB_MOS1_FIELDS(MOD, INS)
#undef MOD
#undef INS
  *flags = (in.IC_GIVEN ? 1 : 0) | (in.OFF ? 2 : 0);
  const int g = c->n;
  const int l[6] = {in.li_Drain, in.li_Gate, in.li_Source, in.li_Bulk, in.li_DrainPrime, in.li_SourcePrime};
  for (int i = 0; i < 6; ++i) lids6[i] = (l[i] == g) ? -1 : l[i];
  return k;
}
int xref_bjt_export(void *h, int idx, double *rec, int *flags, int *lids7) {
  Ctx *c = (Ctx *)h;
  BJT::Instance &in = *static_cast<BJT::Instance *>(c->insts[idx].inst);
  BJT::Model &mo = in.model_;
  int k = 0;
#define MOD(n) rec[k++] = mo.n;
#define INS(n) rec[k++] = in.n;
  XB_BJT_FIELDS(MOD, INS)
#undef MOD
#undef INS
  *flags = (in.IC_GIVEN ? 1 : 0) | (in.OFF ? 2 : 0);
  const int g = c->n;
  const int l[7] = {in.li_Coll, in.li_Base, in.li_Emit, in.li_Subst, in.li_CollP, in.li_BaseP, in.li_EmitP};
  for (int i = 0; i < 7; ++i) lids7[i] = (l[i] == g) ? -1 : l[i];
  return k;
}

void xref_b4_mid(void *h, int idx, double *mid_d, int *mid_i) {
  Ctx *c = (Ctx *)h;
  MOSFET_B4::Instance &in = *static_cast<MOSFET_B4::Instance *>(c->insts[idx].inst);
  int k = 0;
#define PUT(n) mid_d[k++] = in.n;
  XB_B4_MID_REF_D(PUT)
#undef PUT
  k = 0;
  mid_i[k++] = in.origFlag;
#define PUT(n) mid_i[k++] = (int)in.n;
  XB_B4_MID_REF_I(PUT)
#undef PUT
}

}  // extern "C" (reopened below)

// ---- transient run around the REFERENCE device code and the reference tree's Kundert Sparse ----
// Same control flow (tran_driver.h) as the product, but every load is the reference's
// Master::updateState/loadDAEVectors/loadDAEMatrices and every linear solve is ksparse
// (spOrderAndFactor each Newton iteration, as N_LAS_KSparseSolver does).
extern "C" {
char *spCreate(int, int, int *);
double *spGetElement(char *, int, int);
int spOrderAndFactor(char *, double *, double, double, int, int);
int spSolve(char *, double *, double *, double *, double *);
void spClear(char *);
void spDestroy(char *);
}
namespace {
struct Coo { std::vector<int> r, c; std::vector<double> v; };
struct RefBackend {
  Ctx *c;
  int n_;
  std::vector<std::vector<double>> v;
  std::vector<double> J;
  Coo G, C;
  std::vector<int> Gpos, Cpos;
  struct Src { int row; double scale; int type; double p[7]; };
  std::vector<Src> sources;
  std::vector<double> pwl;      // (time, value) pairs of the PWL sources
  void breakpoints(double t, std::vector<double> &out) { for (const Src &q : sources) xb::sim::source_breakpoints(q.type, q.p, pwl.data(), t, out); }
  double max_source_step(double t) {
    double m = 1.0e99;
    for (const Src &q : sources) { const double v = xb::sim::source_max_step(q.type, q.p, t); if (v > 0.0) m = std::min(m, v); }
    return m;
  }
  std::vector<int> probes;
  std::vector<double> times, wave;
  char *M = 0;
  std::vector<double *> addr;
  bool first = true;

  int n() const { return n_; }
  void copy(int d, int a) { v[d] = v[a]; }
  void fill(int d, double x) { std::fill(v[d].begin(), v[d].end(), x); }
  void scale(int d, double a) { for (double &x : v[d]) x = a * x + 0.0 * x; }
  void axpby(int d, double a, int x, double b, int y) { for (int i = 0; i < n_; ++i) v[d][i] = a * v[x][i] + b * v[y][i]; }
  void axpy(int d, double a, int x) { for (int i = 0; i < n_; ++i) v[d][i] = 1.0 * v[d][i] + a * v[x][i]; }
  double norm2(int x) { double s = 0; for (double t : v[x]) s += t * t; return std::sqrt(s); }
  double norm_inf(int x) { double s = 0; for (double t : v[x]) s = std::max(s, std::fabs(t)); return s; }
  double wmax_norm(int x, int w) { double s = 0; for (int i = 0; i < n_; ++i) s = std::max(s, std::fabs(v[x][i] / v[w][i])); return s; }
  double wrms_norm(int x, int w) { double s = 0; for (int i = 0; i < n_; ++i) { double t = v[x][i] / v[w][i]; s += t * t; } return std::sqrt(s / n_); }
  void sol_weights(int d, double rel, double ab, int a, int b) { for (int i = 0; i < n_; ++i) v[d][i] = rel * std::max(std::fabs(v[a][i]), std::fabs(v[b][i])) + ab; }
  void abs_weights(int d, double rel, double ab, int a) { for (int i = 0; i < n_; ++i) v[d][i] = rel * std::fabs(v[a][i]) + ab; }

  bool load_rhs(const xb::sim::Flags &fl, double time) {
    SolverState &s = c->solState;
    s.dcopFlag = fl.dcop; s.tranopFlag = fl.tranop; s.transientFlag = fl.transient; s.initTranFlag_ = fl.initTran;
    s.newtonIter = fl.newtonIter; s.initJctFlag_ = fl.initJct; s.initFixFlag = fl.initFix; s.currTimeStep_ = fl.currTimeStep;
    s.lastTimeStep_ = fl.lastTimeStep; s.beginIntegrationFlag_ = fl.beginIntegration != 0;
    std::copy(v[xb::sim::vNextSol].begin(), v[xb::sim::vNextSol].end(), c->sol.begin());
    c->sol[c->n] = 0.0;
    for (auto *q : {&c->f, &c->q, &c->b, &c->fl, &c->ql}) std::fill(q->begin(), q->end(), 0.0);
    bool ok = c->updateAll();
    ok = c->loadVectorsAll() && ok;
    for (size_t k = 0; k < G.r.size(); ++k) c->f[G.r[k]] += G.v[k] * c->sol[G.c[k]];
    for (size_t k = 0; k < C.r.size(); ++k) c->q[C.r[k]] += C.v[k] * c->sol[C.c[k]];
    for (const Src &q : sources) c->b[q.row] += q.scale * xb::sim::source_value(q.type, q.p, time, pwl.data(), fl.bpTol);
    std::copy(c->f.begin(), c->f.begin() + n_, v[xb::sim::vF].begin());
    std::copy(c->q.begin(), c->q.begin() + n_, v[xb::sim::vQ].begin());
    std::copy(c->b.begin(), c->b.begin() + n_, v[xb::sim::vB].begin());
    std::copy(c->fl.begin(), c->fl.begin() + n_, v[xb::sim::vFlim].begin());
    std::copy(c->ql.begin(), c->ql.begin() + n_, v[xb::sim::vQlim].begin());
    return ok;
  }
  void load_jacobian(double qs, double fs) {
    c->dFdx.put(0.0); c->dQdx.put(0.0);
    c->loadMatricesAll();
    for (size_t k = 0; k < G.r.size(); ++k) c->dFdx.vals[Gpos[k]] += G.v[k];
    for (size_t k = 0; k < C.r.size(); ++k) c->dQdx.vals[Cpos[k]] += C.v[k];
    for (size_t k = 0; k < J.size(); ++k) J[k] = qs * c->dQdx.vals[k] + fs * c->dFdx.vals[k];
  }
  int solve() {
    const CsrMatrix &A = c->dFdx;
    if (!M) {
      int err = 0;
      M = spCreate(n_, 0, &err);
      for (int i = 0; i < n_; ++i)
        for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) addr.push_back(spGetElement(M, i + 1, A.colind[k] + 1));
    } else {
      spClear(M);
    }
    for (size_t k = 0; k < J.size(); ++k) *addr[k] = J[k];
    int rc = spOrderAndFactor(M, 0, 1.0e-3, 1.0e-13, 1, first ? 1 : 0);
    first = false;
    if (rc) { fill(xb::sim::vDX, 0.0); return rc; }
    std::vector<double> b(v[xb::sim::vRHS]);
    return spSolve(M, b.data() - 1, v[xb::sim::vDX].data() - 1, 0, 0);
  }
  bool all_devices_converged() {
    bool all = true;
    for (auto &r : c->insts) all = all && r.inst->isConverged();
    return all;
  }
  void residual_and_norms(const xb::sim::ResidualForm &f, xb::sim::NewtonNorms &out) {
    using namespace xb::sim;
    if (f.form == 0) {            // OneStep::obtainResidual (N_TIA_OneStep.C:219-281)
      axpby(vRHS, 1.0, vQ, -1.0, vQh0);
      axpby(vTmp, f.fs, vF, -f.fs, vB);
      axpby(vRHS, f.inv_h, vRHS, 1.0, vTmp);
      if (f.order2) axpy(vRHS, 0.5, vQh2);
      scale(vRHS, -1.0);
      if (f.limiter) { axpy(vRHS, f.qlim_coef, vQlim); axpy(vRHS, f.fs, vFlim); }
    } else if (f.form == 1) {     // Gear12::obtainResidual (N_TIA_Gear12.C:208-262)
      axpby(vRHS, f.a0, vQ, f.a1, vQh0);
      if (f.order2) axpy(vRHS, f.a2, vQh1);
      axpby(vTmp, 1.0, vF, -1.0, vB);
      axpby(vRHS, f.inv_h, vRHS, 1.0, vTmp);
      scale(vRHS, -1.0);
      if (f.limiter) { axpy(vRHS, f.qlim_coef, vQlim); axpy(vRHS, 1.0, vFlim); }
    } else {                      // NoTimeIntegration::obtainResidual (N_TIA_NoTimeIntegration.C:161-173)
      axpby(vRHS, 1.0, vF, -1.0, vB);
      scale(vRHS, -1.0);
      if (f.limiter) axpy(vRHS, 1.0, vFlim);
    }
    out.rhs_norm2 = norm2(vRHS);
    out.rhs_norm_inf = norm_inf(vRHS);
    out.dx_wmax = wmax_norm(vDX, vSolWt);
    out.devices_converged = all_devices_converged();
  }
  bool limiter_active() const { return c->devOptions.voltageLimiterFlag; }
  void accept_state() { c->currSta = c->nextSta; c->lastSto = c->currSto; c->currSto = c->nextSto; }
  void record(double t) { times.push_back(t); for (int p : probes) wave.push_back(v[xb::sim::vNextSol][p]); }
};
}  // namespace

// step sequence for the next xref_tran_run (TranParams::replay_h / replay_order); consumed by that run
static int g_pwl_n = 0;
static const double *g_pwl = nullptr;
static int g_replay_n = 0;
static const double *g_replay_h = nullptr;
static const int *g_replay_order = nullptr;

extern "C" {

void xref_tran_pwl(int n_points, const double *tv_pairs) { g_pwl_n = n_points; g_pwl = tv_pairs; }      // PWL table of the next xref_tran_run
void xref_tran_replay(int n, const double *h, const int *order) { g_replay_n = n; g_replay_h = h; g_replay_order = order; }

// params: tstop, tstep, delmax, method (0 / 7 trapezoid, 8 Gear), dcop (0 / 1).  Linear part as COO (G, C), sources {row, scale, type, p[7]}.
int xref_tran_run(void *h, const double *params5, const double *x0, int nG, const int *gr, const int *gc, const double *gv,
                  int nC, const int *cr, const int *cc, const double *cv, int ns, const int *srow, const double *sscale,
                  const int *stype, const double *sp7, int n_probes, const int *probes, int max_out, int *n_out,
                  double *times, double *wave, int max_steps, int *n_steps, double *step_info5, double *stats16) {
  Ctx *c = (Ctx *)h;
  RefBackend B;
  B.c = c; B.n_ = c->n;
  B.v.assign(xb::sim::kNumVec, std::vector<double>(c->n, 0.0));
  B.J.assign(c->dFdx.vals.size(), 0.0);
  auto pos = [&](int r, int col) {
    const int *b = &c->dFdx.colind[c->dFdx.rowptr[r]], *e = &c->dFdx.colind[c->dFdx.rowptr[r + 1]];
    return (int)(std::lower_bound(b, e, col) - c->dFdx.colind.data());
  };
  for (int k = 0; k < nG; ++k) if (gr[k] >= 0 && gc[k] >= 0) { B.G.r.push_back(gr[k]); B.G.c.push_back(gc[k]); B.G.v.push_back(gv[k]); B.Gpos.push_back(pos(gr[k], gc[k])); }
  for (int k = 0; k < nC; ++k) if (cr[k] >= 0 && cc[k] >= 0) { B.C.r.push_back(cr[k]); B.C.c.push_back(cc[k]); B.C.v.push_back(cv[k]); B.Cpos.push_back(pos(cr[k], cc[k])); }
  for (int k = 0; k < ns; ++k) if (srow[k] >= 0) { RefBackend::Src q; q.row = srow[k]; q.scale = sscale[k]; q.type = stype[k]; std::memcpy(q.p, sp7 + 7 * k, 7 * sizeof(double)); B.sources.push_back(q); }
  B.probes.assign(probes, probes + n_probes);
  std::copy(x0, x0 + c->n, B.v[xb::sim::vNextSol].begin());
  std::copy(x0, x0 + c->n, B.v[xb::sim::vCurrSol].begin());
  c->solState.transientFlag = true;
  xb::sim::TranParams P;
  P.tstop = params5[0]; P.tstep = params5[1]; P.delmax = params5[2];
  if ((int)params5[3] == 8) P.method = 8;
  P.dcop = params5[4] != 0.0;
  if (g_pwl_n > 0) B.pwl.assign(g_pwl, g_pwl + 2 * (size_t)g_pwl_n);
  g_pwl_n = 0;
  if (g_replay_n > 0) { P.replay_h.assign(g_replay_h, g_replay_h + g_replay_n); P.replay_order.assign(g_replay_order, g_replay_order + g_replay_n); }
  g_replay_n = 0;
  xb::sim::TransientDriver<RefBackend> drv(B, P);
  const int rc = drv.run();
  *n_out = std::min((int)B.times.size(), max_out);
  for (int i = 0; i < *n_out; ++i) { times[i] = B.times[i]; for (int p = 0; p < n_probes; ++p) wave[(size_t)i * n_probes + p] = B.wave[(size_t)i * n_probes + p]; }
  *n_steps = std::min((int)drv.steps.size(), max_steps);
  for (int i = 0; i < *n_steps; ++i) { const auto &r = drv.steps[i]; double *o = step_info5 + 5 * (size_t)i; o[0] = r.t; o[1] = r.h; o[2] = r.newton_iters; o[3] = r.order; o[4] = r.status; }
  const auto &t = drv.stats;
  const double st[16] = {(double)t.accepted, (double)t.rejected, (double)t.newton_total, (double)t.jacobian_loads, (double)t.residual_loads, (double)t.linear_solves, 0, 0, (double)B.times.size(), (double)drv.steps.size(), (double)rc, (double)t.dcop_newton, (double)t.dcop_status, 0, 0, 0};
  std::memcpy(stats16, st, sizeof(st));
  if (B.M) spDestroy(B.M);
  return rc;
}

// The pattern must also hold the linear-device entries: call before xref_finalize.
void xref_add_pattern_entries(void *h, int n, const int *r, const int *cidx) {
  Ctx *c = (Ctx *)h;
  for (int k = 0; k < n; ++k) if (r[k] >= 0 && cidx[k] >= 0) c->extra_pattern.push_back(std::make_pair(r[k], cidx[k]));
}

// ---- Kundert Sparse 1.3 (reference tree: LinearAlgebraServicesPKG/ksparse) ----
// Same call sequence as Epetra_CrsKundertSparse (ksparse/Epetra_CrsKundertSparse.C:60-190):
// spCreate, spGetElement per CSR entry (1-based), spOrderAndFactor, spSolve.
int xref_ksparse_solve(int n, const int *rowptr, const int *colind, const double *vals, const double *rhs, double *x) {
  int err = 0;
  char *M = spCreate(n, 0, &err);
  if (err) return 100 + err;
  for (int i = 0; i < n; ++i)
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) *spGetElement(M, i + 1, colind[k] + 1) = vals[k];
  // thresholds: N_LAS_KSparseSolver defaults (relative 1e-3, absolute 1e-13, diagonal pivoting on)
  int rc = spOrderAndFactor(M, 0, 1.0e-3, 1.0e-13, 1, 1);
  if (!rc) {
    std::vector<double> b(rhs, rhs + n), s(n);
    rc = spSolve(M, b.data() - 1, s.data() - 1, 0, 0);
    std::copy(s.begin(), s.end(), x);
  }
  spDestroy(M);
  return rc;
}

}  // extern "C"
