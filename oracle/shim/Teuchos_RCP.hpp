// oracle build only: minimal stand-in for Teuchos::RCP so that the reference headers
// that merely *mention* reference-counted pointers (N_UTL_Expression.h:50,
// N_DEV_DeviceMgr.h:41) parse.  No Trilinos code is reproduced; this is std::shared_ptr.
#ifndef XB_ORACLE_TEUCHOS_RCP_SHIM
#define XB_ORACLE_TEUCHOS_RCP_SHIM
#include <memory>
namespace Teuchos {
template <class T> class RCP {
 public:
  RCP() {}
  explicit RCP(T *p) : p_(p) {}
  template <class U> RCP(const RCP<U> &o) : p_(o.shared()) {}
  T *operator->() const { return p_.get(); }
  T &operator*() const { return *p_; }
  T *get() const { return p_.get(); }
  bool is_null() const { return !p_; }
  const std::shared_ptr<T> &shared() const { return p_; }
 private:
  std::shared_ptr<T> p_;
};
template <class T> RCP<T> rcp(T *p) { return RCP<T>(p); }
template <class T> bool is_null(const RCP<T> &p) { return p.is_null(); }
template <class T, class U> RCP<T> rcp_dynamic_cast(const RCP<U> &p) {
  RCP<T> r; (void)p; return r;
}
}  // namespace Teuchos
#endif
