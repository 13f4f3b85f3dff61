// oracle build only: placeholder so reference headers that mention
// Teuchos::SerialDenseMatrix (N_UTL_Op.h:88) parse.  Never instantiated by the harness.
#ifndef XB_ORACLE_TEUCHOS_SDM_SHIM
#define XB_ORACLE_TEUCHOS_SDM_SHIM
#include <vector>
namespace Teuchos {
template <class O, class S> class SerialDenseMatrix {
 public:
  SerialDenseMatrix() : r_(0), c_(0) {}
  SerialDenseMatrix(O r, O c) : r_(r), c_(c), v_(r * c) {}
  O numRows() const { return r_; }
  O numCols() const { return c_; }
  S &operator()(O i, O j) { return v_[i + j * r_]; }
  const S &operator()(O i, O j) const { return v_[i + j * r_]; }
  int shape(O r, O c) { r_ = r; c_ = c; v_.assign(r * c, S()); return 0; }
  int reshape(O r, O c) { return shape(r, c); }
  int putScalar(const S &s = S()) { for (auto &x : v_) x = s; return 0; }
 private:
  O r_, c_;
  std::vector<S> v_;
};
}  // namespace Teuchos
#endif
