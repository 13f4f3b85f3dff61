// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Link-time shim that lets the reference's own device-model translation units
// (compiled where they lie under /root/reference, see oracle/Makefile) run without
// Trilinos, the expression engine, the parser or the rest of Xyce.  It supplies
// minimal definitions for the handful of reference classes whose .C files cannot
// be compiled here (they include Teuchos headers): DeviceEntity's parameter
// plumbing, SolverState construction, Configuration registration hooks,
// Util::Param value access and message/diagnostic sinks.  Class declarations come
// from the reference headers; only member *definitions* are provided here, with
// behaviour restricted to what numeric-constant netlists need (no expressions).
//   DeviceEntity::setParams  follows Core/N_DEV_DeviceEntity.C:1979-2240
//   DeviceEntity::given      follows Core/N_DEV_DeviceEntity.C:1922-1929
#include <Xyce_config.h>
#include <iostream>
#include <sstream>
#include <typeinfo>
#include <N_DEV_fwd.h>
#include <N_DEV_Const.h>
#include <N_DEV_DeviceEntity.h>
#include <N_DEV_DeviceOptions.h>
#include <N_DEV_SolverState.h>
#include <N_DEV_Configuration.h>
#include <N_DEV_Message.h>
#include <N_DEV_Param.h>
#include <N_DEV_Pars.h>
#include <N_DEV_Device.h>
#include <N_DEV_DeviceInstance.h>
#include <N_DEV_DeviceModel.h>
#include <N_DEV_InstanceName.h>
#include <N_DEV_NumericalJacobian.h>
#include <N_UTL_Param.h>
#include <N_UTL_Expression.h>
#include <N_DEV_MatrixLoadData.h>
#include <N_UTL_Pack.h>
#include <N_ERH_Message.h>
#include <N_UTL_Demangle.h>
#include <N_UTL_DeviceNameConverters.h>

namespace Xyce {

std::string demangle(const char *symbol) { return std::string(symbol); }

namespace Util {

const char *separator = "------------------------------------------------------------";

// --- Util::Param value access (numeric / string / bool constants only) ---
template <> double Param::getImmutableValue<double>() const {
  switch (data_->enumType()) {
    case DBLE: return getValue<double>();
    case INT:  return getValue<int>();
    case LNG:  return getValue<long>();
    case BOOL: return getValue<bool>() ? 1.0 : 0.0;
    case STR: { std::istringstream is(getValue<std::string>()); double d = 0; is >> d; return d; }
    default: return 0.0;
  }
}
template <> int Param::getImmutableValue<int>() const {
  switch (data_->enumType()) {
    case DBLE: return static_cast<int>(getValue<double>());
    case INT:  return getValue<int>();
    case LNG:  return static_cast<int>(getValue<long>());
    case BOOL: return getValue<bool>() ? 1 : 0;
    case STR: { std::istringstream is(getValue<std::string>()); int d = 0; is >> d; return d; }
    default: return 0;
  }
}
template <> long Param::getImmutableValue<long>() const { return getImmutableValue<int>(); }
template <> bool Param::getImmutableValue<bool>() const { return getImmutableValue<double>() != 0.0; }

std::string Param::stringValue() const {
  if (!data_) return std::string();
  if (data_->enumType() == STR) return getValue<std::string>();
  std::ostringstream os; os << getImmutableValue<double>(); return os.str();
}
std::string Param::uTag() const {
  std::string s(tag_);
  for (auto &c : s) c = std::toupper(c);
  return s;
}
void Param::setTimeDependent(bool) {}
bool Param::isTimeDependent() const { return false; }

// Expression objects are never created by the oracle harness (numeric netlists only);
// these exist so std::vector<Expression> members of SolverState can be destroyed.
Expression::Expression(const Expression &) {}
Expression &Expression::operator=(const Expression &) { return *this; }
Expression::~Expression() {}

void addSymbol(SymbolTable &, SymbolType, int, const std::string &) {}
std::string xyceDeviceNameToSpiceName(std::string &name) { return name; }

}  // namespace Util

template <> int Pack<Util::Param>::packedByteCount(const Util::Param &) { return 0; }
template <> void Pack<Util::Param>::pack(const Util::Param &, char *, int, int &, Parallel::Communicator *) {}
template <> void Pack<Util::Param>::unpack(Util::Param &, char *, int, int &, Parallel::Communicator *) {}

namespace Device {

// ---------------- SolverState ----------------
SolverState::SolverState()
  : isPDESystem_(false), pdt_(0.0), currentOrder_(0), usedOrder_(0), integrationMethod_(0),
    currTimeStep_(0.0), lastTimeStep_(0.0), oldeTimeStep_(0.0), currTime_(0.0), finalTime_(0.0),
    startingTimeStep_(0.0), bpTol_(0.0), acceptedTime_(0.0), mpdeOnFlag_(false), currFastTime_(0.0),
    blockAnalysisFlag_(false), spAnalysisFlag_(false), earlyNoiseFlag_(false),
    doubleDCOPEnabled(false), doubleDCOPStep(0), timeStepNumber_(0),
    ltraDevices_(false), ltraTimeIndex_(0), ltraTimeHistorySize_(0), ltraDoCompact_(false),
    newtonIter(0), continuationStepNumber(0), firstContinuationParam(false), firstSolveComplete(false),
    initTranFlag_(false), beginIntegrationFlag_(false), dcopFlag(false), inputOPFlag(false),
    transientFlag(false), dcsweepFlag(false), tranopFlag(false), acopFlag(false), noiseFlag(false),
    locaEnabledFlag(false), externalInitJctFlag_(false), externalStateFlag_(false),
    initJctFlag_(false), initFixFlag(false), sweepSourceResetFlag(false), debugTimeFlag(false),
    twoLevelNewtonCouplingMode(Nonlinear::FULL_PROBLEM), pdeAlpha_(0.0), PDEcontinuationFlag_(false),
    chargeHomotopy_(false), chargeAlpha_(0.0), artParameterFlag_(false), gainScale_(1.0),
    nltermScale_(1.0), sizeParameterFlag_(false), sizeScale_(1.0), currFreq_(0.0), groupWrapperPtr_(0)
{}
SolverState::~SolverState() {}

// ---------------- Configuration hooks ----------------
void Configuration::addDevice(const char *, const int, ModelTypeId, ModelTypeId, int, int) {}
void Configuration::addModel(const char *, const int, ModelTypeId, ModelTypeId) {}

// ---------------- DeviceEntity ----------------
DeviceEntity::DeviceEntity(ParametricData<void> &parametric_data, const SolverState &solver_state,
                           const DeviceOptions &device_options, const NetlistLocation &netlist_location)
  : defaultParamName_(), parametricData_(parametric_data), netlistLocation_(netlist_location),
    solState_(solver_state), globals_(solver_state.getGlobals()), devOptions_(device_options)
{}
DeviceEntity::~DeviceEntity() {}
void DeviceEntity::processSuccessfulTimeStep() {}
bool DeviceEntity::analyticSensitivityAvailable(const std::string &) { return false; }
bool DeviceEntity::updateDependentParameters() { return true; }
void DeviceEntity::applyDepSolnLIDs() {}
void DeviceEntity::checkParamVersions(double) const {}

bool DeviceEntity::given(const std::string &parameter_name) const {
  ParameterMap::const_iterator it = getParameterMap().find(parameter_name);
  if (it == getParameterMap().end()) {
    std::cerr << "xyce_shim: unrecognized parameter " << parameter_name << std::endl;
    return false;
  }
  return Xyce::Device::wasValueGiven(*this, (*it).second->getSerialNumber());
}

bool DeviceEntity::getParam(const std::string &name, double &result) const {
  ParameterMap::const_iterator it = getParameterMap().find(name);
  if (it == getParameterMap().end()) return false;
  const Descriptor &d = *(*it).second;
  if (d.isType<double>()) result = d.value<double>(*this);
  else if (d.isType<int>()) result = d.value<int>(*this);
  else if (d.isType<bool>()) result = d.value<bool>(*this) ? 1.0 : 0.0;
  else return false;
  return true;
}

bool DeviceEntity::setParam(const std::string &name, double val, bool, bool) {
  ParameterMap::const_iterator it = getParameterMap().find(name);
  if (it == getParameterMap().end()) return false;
  const Descriptor &d = *(*it).second;
  if (d.isType<double>()) d.value<double>(*this) = val;
  else if (d.isType<int>()) d.value<int>(*this) = static_cast<int>(val);
  else if (d.isType<bool>()) d.value<bool>(*this) = (val != 0.0);
  else return false;
  return true;
}
bool DeviceEntity::setDefaultParam(double val, bool o, bool i) { return setParam(defaultParamName_, val, o, i); }

void DeviceEntity::setParams(const std::vector<Param> &params) {
  for (std::vector<Param>::const_iterator p = params.begin(); p != params.end(); ++p) {
    Param &param = const_cast<Param &>(*p);
    ParameterMap::const_iterator eit = getParameterMap().find(param.tag());
    if (eit == getParameterMap().end()) {
      std::cerr << "xyce_shim: parameter " << param.tag() << " not known to this entity" << std::endl;
      continue;
    }
    const Descriptor &descriptor = *(*eit).second;
    if (descriptor.hasGivenMember()) {
      if (param.given()) descriptor.setGiven(*this, true);
      else if (descriptor.getGiven(*this)) continue;
    }
    Xyce::Device::setValueGiven(*this, descriptor.getSerialNumber(), param.given());
    if (param.given() || param.default_val()) {
      if (descriptor.isType<double>()) {
        descriptor.value<double>(*this) = param.getImmutableValue<double>();
        if (isTempParam(param.tag()) && descriptor.getAutoConvertTemperature())
          descriptor.value<double>(*this) += CONSTCtoK;
        if (descriptor.hasOriginalValueStored())
          Xyce::Device::setOriginalValue(*this, descriptor.getSerialNumber(), descriptor.value<double>(*this));
      } else if (descriptor.isType<std::string>()) {
        descriptor.value<std::string>(*this) = param.stringValue();
      } else if (descriptor.isType<int>()) {
        descriptor.value<int>(*this) = param.getImmutableValue<int>();
      } else if (descriptor.isType<long>()) {
        descriptor.value<long>(*this) = param.getImmutableValue<long>();
      } else if (descriptor.isType<bool>()) {
        descriptor.value<bool>(*this) = param.getImmutableValue<bool>();
      }
    }
  }
}

// ---------------- diagnostics the device code references ----------------
void could_not_find_model_error(const Device &, const std::string &m, const std::string &i, const NetlistLocation &)
{ std::cerr << "xyce_shim: could not find model " << m << " for " << i << std::endl; }
void duplicate_entity_warning(const Device &, const DeviceEntity &, const NetlistLocation &) {}
void duplicate_instance_warning(const Device &, const DeviceInstance &, const NetlistLocation &) {}
void duplicate_model_warning(const Device &, const DeviceModel &, const NetlistLocation &) {}
void instance_must_reference_model_error(const Device &, const std::string &m, const NetlistLocation &)
{ std::cerr << "xyce_shim: instance must reference model " << m << std::endl; }

NumericalJacobian::NumericalJacobian(MatrixLoadData &mlData1, const SolverState &ss1, const ExternData &ed1, const DeviceOptions &do1)
  : mlData(mlData1), cols(mlData1.cols), vals(mlData1.vals), Qvals(mlData1.Qvals),
    val_local(mlData1.val_local), Qval_local(mlData1.Qval_local), col_local(mlData1.col_local),
    row_local(mlData1.row_local), internalFlag(mlData1.internalFlag), devOptions(do1), solState(ss1),
    extData(ed1), maxCols(10)
{}
NumericalJacobian::~NumericalJacobian() {}
bool NumericalJacobian::testDAEMatrices(DeviceInstance &, const std::vector<const std::string *> &) { return true; }
void NumericalJacobian::loadLocalDAEVectorsIncludingB(DeviceInstance &) {}

}  // namespace Device

namespace Report {
unsigned get_message_count(unsigned) { return 0; }
void report_message(const char *message, unsigned, const MessageCode &) { std::cerr << "xyce_ref: " << message << std::endl; }
}  // namespace Report

}  // namespace Xyce
