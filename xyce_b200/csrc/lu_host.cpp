// xyce_b200 -- sparse direct solver, host side: the symbolic phase and the first (pivoting)
// numeric factorization that fix the pattern and pivot sequence the GPU kernels then reuse.
//
// What it stands in for: Amesos_Klu::SymbolicFactorization / NumericFactorization as driven by
// Linear::AmesosSolver::doSolve (src/LinearAlgebraServicesPKG/N_LAS_AmesosSolver.C:216-470).
// The arithmetic of that path lives in SuiteSparse KLU/BTF/AMD inside Trilinos Amesos, which is
// NOT part of the reference tree.  This file restates KLU's published algorithm
// (Davis & Palamadai Natarajan, ACM TOMS 37(3), 2010) from the description, not from source:
//   1. maximum transversal  -> zero-free diagonal                     (BTF "maxtrans")
//   2. strongly connected components -> upper block triangular form   (BTF "strongcomp")
//   3. fill-reducing ordering of each diagonal block (minimum degree on A+A')   (role of AMD)
//   4. Gilbert-Peierls left-looking LU of each block with threshold partial pivoting that
//      prefers the diagonal (tol = 0.001, KLU default); off-diagonal blocks are kept unfactored
//      and used by block back-substitution.
// The result (LuPlan) is what xgpu_lu_refactor / xgpu_lu_solve consume: klu_refactor semantics
// (pattern and pivot order reused, "Refactorize = !KLU_REPIVOT", N_LAS_AmesosSolver.C:316-318).
#include "lu.h"

#include <algorithm>
#include <cmath>
#include <map>
#include <numeric>
#include <queue>

namespace xb {
namespace lu {

namespace {

// ---- 1. maximum transversal (augmenting-path DFS with cheap assignment) ----
// Input: CSC pattern (column j has rows Ai[Ap[j]..Ap[j+1])).  Output: match[i] = column matched to row i.
int max_transversal(int n, const std::vector<int> &Ap, const std::vector<int> &Ai, std::vector<int> &col_of_row) {
  col_of_row.assign(n, -1);
  std::vector<int> cheap(Ap.begin(), Ap.end() - 1), seen(n, -1), js(n), is(n), ps(n);
  int nmatch = 0;
  for (int k = 0; k < n; ++k) {
    int head = 0;
    bool found = false;
    js[0] = k;
    while (head >= 0) {
      const int j = js[head];
      if (seen[j] != k) {               // first visit of column j in this search
        seen[j] = k;
        int p = cheap[j];
        for (; p < Ap[j + 1] && !found; ++p) {
          const int i = Ai[p];
          if (col_of_row[i] == -1) { found = true; is[head] = i; }
        }
        cheap[j] = p;
        if (found) break;
        ps[head] = Ap[j];
      }
      int p = ps[head];
      for (; p < Ap[j + 1]; ++p) {
        const int i = Ai[p];
        if (seen[col_of_row[i]] == k) continue;
        ps[head] = p + 1;
        is[head] = i;
        js[++head] = col_of_row[i];
        break;
      }
      if (p == Ap[j + 1]) --head;
    }
    if (found) {
      for (int h = head; h >= 0; --h) col_of_row[is[h]] = js[h];
      ++nmatch;
    }
  }
  return nmatch;
}

// ---- 2. strongly connected components (iterative Tarjan) of the graph with edge j -> i for
//         every entry (i, j) of the column-permuted matrix (rows already aligned to the diagonal) ----
// Produces blocks in an order that makes the permuted matrix UPPER block triangular.
void strong_components(int n, const std::vector<int> &Ap, const std::vector<int> &Ai, std::vector<int> &perm,
                       std::vector<int> &block_ptr) {
  std::vector<int> index(n, -1), low(n, 0), onstack(n, 0), st, callstack, pos(n, 0);
  perm.clear();
  block_ptr.assign(1, 0);
  int counter = 0;
  for (int s = 0; s < n; ++s) {
    if (index[s] != -1) continue;
    callstack.push_back(s);
    index[s] = low[s] = counter++;
    st.push_back(s); onstack[s] = 1;
    pos[s] = Ap[s];
    while (!callstack.empty()) {
      const int v = callstack.back();
      if (pos[v] < Ap[v + 1]) {
        const int w = Ai[pos[v]++];
        if (index[w] == -1) {
          index[w] = low[w] = counter++;
          st.push_back(w); onstack[w] = 1;
          pos[w] = Ap[w];
          callstack.push_back(w);
        } else if (onstack[w]) {
          low[v] = std::min(low[v], index[w]);
        }
      } else {
        callstack.pop_back();
        if (!callstack.empty()) low[callstack.back()] = std::min(low[callstack.back()], low[v]);
        if (low[v] == index[v]) {
          int w;
          do {
            w = st.back(); st.pop_back(); onstack[w] = 0;
            perm.push_back(w);
          } while (w != v);
          block_ptr.push_back((int)perm.size());
        }
      }
    }
  }
  // Tarjan emits a component after everything reachable from it.  With edges column -> row
  // (j -> i for a(i,j)), a component that is reached (rows) comes first; entries a(i,j) then have
  // block(i) <= block(j): upper block triangular.
}

// ---- 3. minimum-degree ordering on the pattern of B + B' (B = one diagonal block) ----
// A plain quotient-free minimum degree with lazy priority queue: adequate for the block sizes of
// circuit matrices after BTF; plays the role AMD plays inside KLU.
void min_degree(int n, const std::vector<std::vector<int>> &adj_in, std::vector<int> &order) {
  std::vector<std::vector<int>> adj(adj_in);
  for (auto &a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
  std::vector<char> done(n, 0);
  // "dense" nodes (supply rails, clock nets: degree > 10 sqrt(n), the rule AMD uses) are taken out of the
  // graph and ordered last; without this every elimination next to a rail merges into an O(n) list
  std::vector<int> dense;
  {
    const double thr = std::max(16.0, 10.0 * std::sqrt((double)n));
    for (int i = 0; i < n; ++i) if ((double)adj[i].size() > thr) { dense.push_back(i); done[i] = 1; }
    if (!dense.empty())
      for (int i = 0; i < n; ++i) {
        if (done[i]) { adj[i].clear(); continue; }
        std::vector<int> &a = adj[i];
        a.erase(std::remove_if(a.begin(), a.end(), [&](int w) { return done[w] != 0; }), a.end());
      }
  }
  typedef std::pair<int, int> DI;
  std::priority_queue<DI, std::vector<DI>, std::greater<DI>> pq;
  for (int i = 0; i < n; ++i) if (!done[i]) pq.push(DI((int)adj[i].size(), i));
  order.clear();
  std::vector<int> merged;
  const int n_sparse = n - (int)dense.size();
  while ((int)order.size() < n_sparse) {
    DI top = pq.top(); pq.pop();
    const int v = top.second;
    if (done[v] || top.first != (int)adj[v].size()) continue;
    done[v] = 1;
    order.push_back(v);
    // eliminate v: its neighbours become a clique
    std::vector<int> &nb = adj[v];
    for (int u : nb) {
      std::vector<int> &au = adj[u];
      merged.clear();
      std::set_union(au.begin(), au.end(), nb.begin(), nb.end(), std::back_inserter(merged));
      au.clear();
      for (int w : merged) if (w != u && w != v && !done[w]) au.push_back(w);
      pq.push(DI((int)au.size(), u));
    }
    nb.clear(); nb.shrink_to_fit();
  }
  for (int v : dense) order.push_back(v);
}

// ---- 4. Gilbert-Peierls left-looking LU of one block with threshold partial pivoting ----
// B is given by columns in the block's own (ordered) numbering: col j has (row, value) pairs.
// Output: pinv (row -> pivot position), L/U in CSC over pivot positions; L unit-diagonal (not stored),
// U column holds off-diagonal entries (sorted ascending) followed by the diagonal as its LAST entry.
struct BlockLU {
  int n = 0;
  std::vector<int> Lp, Li, Up, Ui, pinv;
  std::vector<double> Lx, Ux;
  bool singular = false;
};

void factor_block(int n, const std::vector<std::vector<std::pair<int, double>>> &cols, double tol, BlockLU &f) {
  f.n = n;
  f.Lp.assign(1, 0); f.Up.assign(1, 0);
  f.Li.clear(); f.Ui.clear(); f.Lx.clear(); f.Ux.clear();
  f.pinv.assign(n, -1);
  std::vector<double> x(n, 0.0);
  std::vector<int> mark(n, -1), topo, dfs_stack, dfs_pos(n), pattern;
  std::vector<int> prow(n, -1);   // pivot position -> original row
  for (int k = 0; k < n; ++k) {
    // symbolic: reach of the column's rows through the graph of L (over pivoted rows)
    topo.clear(); pattern.clear();
    for (const auto &e : cols[k]) {
      int r = e.first;
      if (mark[r] == k) continue;
      // DFS in pivot-position space for pivoted rows
      dfs_stack.clear();
      dfs_stack.push_back(r);
      mark[r] = k;
      dfs_pos[r] = (f.pinv[r] >= 0) ? f.Lp[f.pinv[r]] : 0;
      while (!dfs_stack.empty()) {
        const int v = dfs_stack.back();
        const int pv = f.pinv[v];
        bool pushed = false;
        if (pv >= 0) {
          while (dfs_pos[v] < f.Lp[pv + 1]) {
            const int w = f.Li[dfs_pos[v]++];   // original row index stored in Li during factorization
            if (mark[w] == k) continue;
            mark[w] = k;
            dfs_pos[w] = (f.pinv[w] >= 0) ? f.Lp[f.pinv[w]] : 0;
            dfs_stack.push_back(w);
            pushed = true;
            break;
          }
        }
        if (!pushed) { dfs_stack.pop_back(); topo.push_back(v); }
      }
    }
    // numeric: x = B(:,k); sparse triangular solve in topological order (reverse post-order)
    for (int v : topo) x[v] = 0.0;
    for (const auto &e : cols[k]) x[e.first] += e.second;
    for (int t = (int)topo.size() - 1; t >= 0; --t) {
      const int v = topo[t];
      const int pv = f.pinv[v];
      if (pv < 0) continue;
      const double xv = x[v];
      for (int p = f.Lp[pv]; p < f.Lp[pv + 1]; ++p) x[f.Li[p]] -= f.Lx[p] * xv;
    }
    // pivot search among non-pivoted rows; prefer the diagonal (row k) if within tol of the largest
    double amax = 0.0; int imax = -1;
    for (int v : topo) if (f.pinv[v] < 0) { const double a = std::fabs(x[v]); if (a > amax) { amax = a; imax = v; } }
    int piv = imax;
    if (mark[k] == k && f.pinv[k] < 0 && std::fabs(x[k]) >= tol * amax && x[k] != 0.0) piv = k;
    if (piv < 0 || x[piv] == 0.0) {
      f.singular = true;
      // keep the structure consistent: choose any unpivoted row (the structural diagonal if free)
      if (piv < 0) { for (int r = 0; r < n; ++r) if (f.pinv[r] < 0) { piv = r; break; } }
    }
    const double pivot = x[piv];
    // U(:,k): pivoted rows (by pivot position), diagonal last
    std::vector<std::pair<int, double>> ucol;
    for (int v : topo) if (f.pinv[v] >= 0) ucol.push_back(std::make_pair(f.pinv[v], x[v]));
    std::sort(ucol.begin(), ucol.end());
    for (auto &u : ucol) { f.Ui.push_back(u.first); f.Ux.push_back(u.second); }
    f.Ui.push_back(k); f.Ux.push_back(pivot);
    f.Up.push_back((int)f.Ui.size());
    // L(:,k): remaining unpivoted rows (original row ids for now), scaled by the pivot
    f.pinv[piv] = k; prow[k] = piv;
    for (int v : topo) if (f.pinv[v] < 0) { f.Li.push_back(v); f.Lx.push_back(pivot != 0.0 ? x[v] / pivot : 0.0); }
    f.Lp.push_back((int)f.Li.size());
  }
  // L row indices: original rows -> pivot positions, sorted within each column
  for (int k = 0; k < n; ++k) {
    std::vector<std::pair<int, double>> lcol;
    for (int p = f.Lp[k]; p < f.Lp[k + 1]; ++p) lcol.push_back(std::make_pair(f.pinv[f.Li[p]], f.Lx[p]));
    std::sort(lcol.begin(), lcol.end());
    for (int p = f.Lp[k], t = 0; p < f.Lp[k + 1]; ++p, ++t) { f.Li[p] = lcol[t].first; f.Lx[p] = lcol[t].second; }
  }
}

}  // namespace

// ---- batched groups: blocks with one common symbolic pattern, compiled into bundle programs (lu.h) ----------------
namespace {
struct BOp { int type, dst, a, b; };

// Greedy list scheduling in program order.  `nres` resources; op reads r[0..nr) and read-modify-writes w.
// A bundle is executed by kBundle lanes at once (one op each): all its loads precede all its stores, so an op may
// share a bundle with an EARLIER op that reads what it writes, never precede it; ops that update the same destination
// stay in program order (RAW chain): summation order per factor entry is exactly the left-looking order of the
// warp-per-block kernels.  Bundles are TYPED (all ops of a bundle have one type: the lanes of a tile do not diverge).
void schedule_bundles(const std::vector<BOp> &ops, int nres, const std::vector<std::vector<int>> &reads, const std::vector<int> &writes,
                      std::vector<unsigned short> &prog, int &nbundles) {
  std::vector<int> lastW(nres, -1), lastR(nres, -1), btype;
  std::vector<std::vector<int>> bundles;
  for (size_t i = 0; i < ops.size(); ++i) {
    int earliest = 0;
    for (int r : reads[i]) earliest = std::max(earliest, lastW[r] + 1);
    if (writes[i] >= 0) earliest = std::max(earliest, std::max(lastW[writes[i]] + 1, lastR[writes[i]]));
    size_t bidx = (size_t)earliest;
    while (bidx < bundles.size() && ((int)bundles[bidx].size() >= kBundle || btype[bidx] != ops[i].type)) ++bidx;
    if (bidx >= bundles.size()) { bundles.resize(bidx + 1); btype.resize(bidx + 1, -1); }
    if (btype[bidx] < 0) btype[bidx] = ops[i].type;
    bundles[bidx].push_back((int)i);
    for (int r : reads[i]) lastR[r] = std::max(lastR[r], (int)bidx);
    if (writes[i] >= 0) lastW[writes[i]] = (int)bidx;
  }
  // bundles left empty by the search (a later op needed a later slot) stay as no-ops
  nbundles = (int)bundles.size();
  prog.assign((size_t)nbundles * kBundle * 4, 0);
  for (int bi = 0; bi < nbundles; ++bi)
    for (int k = 0; k < kBundle; ++k) {
      unsigned short *o = &prog[((size_t)bi * kBundle + k) * 4];      // 8 bytes per op: {dst | type << 14, a, b, 0}
      if (k < (int)bundles[bi].size()) {
        const BOp &op = ops[bundles[bi][k]];
        o[0] = (unsigned short)(op.dst | (op.type << 14)); o[1] = (unsigned short)op.a; o[2] = (unsigned short)op.b;
      } else {
        o[0] = (unsigned short)(kOpNop << 14); o[1] = 0; o[2] = 0;
      }
    }
}

bool g_batching = true;
void build_batch_groups(LuPlan &plan, const std::vector<int> &level) {
  const std::vector<int> &bptr = plan.block_ptr;
  const int nblocks = (int)bptr.size() - 1;
  plan.batch.clear();
  if (!g_batching) return;
  // signature of a block: everything the programs depend on, relative to the block's first position / entry
  std::map<std::vector<int>, std::vector<int>> by_sig;
  for (int b = 0; b < nblocks; ++b) {
    const int k0 = bptr[b], k1 = bptr[b + 1], nb = k1 - k0;
    if (nb < 2 || nb > kBigBlock) continue;
    const int l0 = plan.Lp[k0], u0 = plan.Up[k0], a0 = plan.acol_ptr[k0];
    const int nl = plan.Lp[k1] - l0, nu = plan.Up[k1] - u0, na = plan.acol_ptr[k1] - a0;
    if (nu + nl + nb > kBatchMaxSlots) continue;
    std::vector<int> sig;
    sig.reserve(4 + 3 * (nb + 1) + nl + nu + na);
    sig.push_back(nb); sig.push_back(nl); sig.push_back(nu); sig.push_back(level[b]);
    for (int k = k0; k <= k1; ++k) { sig.push_back(plan.Lp[k] - l0); sig.push_back(plan.Up[k] - u0); sig.push_back(plan.acol_ptr[k] - a0); }
    for (int q = l0; q < l0 + nl; ++q) sig.push_back(plan.Li[q] - k0);
    for (int q = u0; q < u0 + nu; ++q) sig.push_back(plan.Ui[q] - k0);
    for (int q = a0; q < a0 + na; ++q) { const int d = plan.acol_dst[q]; sig.push_back(d >= 0 ? d - u0 : nu + (~d - l0)); }
    by_sig[sig].push_back(b);
  }
  for (auto &kv : by_sig) {
    const std::vector<int> &blocks = kv.second;
    if ((int)blocks.size() < kBatchMinBlocks) continue;
    LuPlan::BatchGroup g;
    const int b0 = blocks[0], k0 = bptr[b0], k1 = bptr[b0 + 1];
    const int l0 = plan.Lp[k0], u0 = plan.Up[k0], a0 = plan.acol_ptr[k0];
    g.nb = k1 - k0; g.nl = plan.Lp[k1] - l0; g.nu = plan.Up[k1] - u0; g.na = plan.acol_ptr[k1] - a0; g.level = level[b0];
    g.blocks = blocks;
    const int nblk = (int)blocks.size();
    auto uslot = [&](int q) { return q - u0; };
    auto lslot = [&](int q) { return g.nu + (q - l0); };
    g.a_dst.resize(g.na);
    for (int e = 0; e < g.na; ++e) { const int d = plan.acol_dst[a0 + e]; g.a_dst[e] = d >= 0 ? uslot(d) : lslot(~d); }
    g.a_src.resize((size_t)g.na * nblk);
    for (int j = 0; j < nblk; ++j) {
      const int aj = plan.acol_ptr[bptr[blocks[j]]];
      for (int e = 0; e < g.na; ++e) g.a_src[(size_t)e * nblk + j] = plan.acol_src[aj + e];
    }
    // ---- refactor program: left-looking column by column, on factor slots ----
    {
      std::vector<BOp> ops; std::vector<std::vector<int>> reads; std::vector<int> writes;
      // slot of row r in column k (U part, pivot or L part)
      auto col_slot = [&](int k, int r) {
        const int ub = plan.Up[k], ue = plan.Up[k + 1] - 1, lb = plan.Lp[k], le = plan.Lp[k + 1];
        if (r == k) return uslot(ue);
        if (r < k) return uslot((int)(std::lower_bound(plan.Ui.begin() + ub, plan.Ui.begin() + ue, r) - plan.Ui.begin()));
        return lslot((int)(std::lower_bound(plan.Li.begin() + lb, plan.Li.begin() + le, r) - plan.Li.begin()));
      };
      for (int k = k0; k < k1; ++k) {
        const int ub = plan.Up[k], ue = plan.Up[k + 1] - 1, lb = plan.Lp[k], le = plan.Lp[k + 1];
        for (int q = ub; q < ue; ++q) {
          const int i = plan.Ui[q];
          for (int t = plan.Lp[i]; t < plan.Lp[i + 1]; ++t) {
            const int dst = col_slot(k, plan.Li[t]);
            ops.push_back({kOpFnma, dst, lslot(t), uslot(q)}); reads.push_back({dst, lslot(t), uslot(q)}); writes.push_back(dst);
          }
        }
        // the division also tests the pivot (zero / non-finite / below threshold); a column without L entries gets an explicit check
        if (lb == le) { ops.push_back({kOpChk, 0, uslot(ue), 0}); reads.push_back({uslot(ue)}); writes.push_back(-1); }
        for (int q = lb; q < le; ++q) { ops.push_back({kOpDiv, lslot(q), uslot(ue), 0}); reads.push_back({lslot(q), uslot(ue)}); writes.push_back(lslot(q)); }
      }
      schedule_bundles(ops, g.nu + g.nl, reads, writes, g.rf_prog, g.rf_bundles);
    }
    // ---- solve program on the block's right-hand side y[0, nb): unit-lower forward, then backward with U ----
    {
      std::vector<BOp> ops; std::vector<std::vector<int>> reads; std::vector<int> writes;
      // resources: y rows only (factor slots are read-only here)
      for (int k = k0; k < k1; ++k)
        for (int q = plan.Lp[k]; q < plan.Lp[k + 1]; ++q) {
          const int dst = plan.Li[q] - k0;
          ops.push_back({kOpFnma, dst, lslot(q), k - k0}); reads.push_back({dst, k - k0}); writes.push_back(dst);
        }
      for (int k = k1 - 1; k >= k0; --k) {
        const int ue = plan.Up[k + 1] - 1;
        ops.push_back({kOpDiv, k - k0, uslot(ue), 0}); reads.push_back({k - k0}); writes.push_back(k - k0);
        for (int q = plan.Up[k]; q < ue; ++q) {
          const int dst = plan.Ui[q] - k0;
          ops.push_back({kOpFnma, dst, uslot(q), k - k0}); reads.push_back({dst, k - k0}); writes.push_back(dst);
        }
      }
      schedule_bundles(ops, g.nb, reads, writes, g.sv_prog, g.sv_bundles);
    }
    for (int b : blocks) plan.block_big[b] = 3;
    plan.batch.push_back(std::move(g));
  }
}
}  // namespace

// Everything the GPU kernels need beyond the permutations and the L / U patterns: scatter maps A -> factor slots,
// off-diagonal entries by column and by row, block levels, pull lists, large-block schedules, flop count.
// Inputs already in the plan: n, block_ptr, row_perm (through new_rowpos), col_perm, Lp / Li / Up / Ui.
// Ap / Ai / Aidx: CSC pattern of A with the CSR value index of every entry; vals may be null (imported plans).
static void finish_plan(LuPlan &plan, const std::vector<int> &Ap, const std::vector<int> &Ai, const std::vector<int> &Aidx,
                        const double *vals, const std::vector<int> &new_rowpos) {
  const int n = plan.n;
  const std::vector<int> &bptr = plan.block_ptr;
  const int nblocks = (int)bptr.size() - 1;
  // scatter maps: every entry of A goes either into a diagonal block column (dense work vector slot)
  // or into the off-diagonal list used by block back-substitution
  plan.acol_ptr.assign(n + 1, 0);
  plan.off_ptr.assign(n + 1, 0);
  std::vector<int> blk_of_pos(n);
  for (int b = 0; b < nblocks; ++b) for (int t = bptr[b]; t < bptr[b + 1]; ++t) blk_of_pos[t] = b;
  for (int t = 0; t < n; ++t) {
    const int j = plan.col_perm[t];
    for (int p = Ap[j]; p < Ap[j + 1]; ++p) {
      const int rp = new_rowpos[Ai[p]];
      if (blk_of_pos[rp] == blk_of_pos[t]) { plan.acol_row.push_back(rp); plan.acol_src.push_back(Aidx[p]); }
      else { plan.off_row.push_back(rp); plan.off_src.push_back(Aidx[p]); plan.off_val_host.push_back(vals ? vals[Aidx[p]] : 0.0); }
    }
    plan.acol_ptr[t + 1] = (int)plan.acol_row.size();
    plan.off_ptr[t + 1] = (int)plan.off_row.size();
  }
  // row-major copy of the off-diagonal entries
  {
    plan.offr_ptr.assign(n + 1, 0);
    for (int r : plan.off_row) ++plan.offr_ptr[r + 1];
    for (int r = 0; r < n; ++r) plan.offr_ptr[r + 1] += plan.offr_ptr[r];
    plan.offr_col.resize(plan.off_row.size()); plan.offr_src.resize(plan.off_row.size());
    std::vector<int> fill(plan.offr_ptr.begin(), plan.offr_ptr.end() - 1);
    for (int t = 0; t < n; ++t)
      for (int q = plan.off_ptr[t]; q < plan.off_ptr[t + 1]; ++q) {
        const int d = fill[plan.off_row[q]]++;
        plan.offr_col[d] = t; plan.offr_src[d] = plan.off_src[q];
      }
  }
  // block dependency levels for the solve: block b needs every block c > b that has an entry in b's rows
  std::vector<int> level(nblocks, 0);
  int maxlevel = 0;
  {
    // off-diagonal entry (row position rp in block b, column position t in block c), c > b
    std::vector<std::vector<int>> dep(nblocks);
    for (int t = 0; t < n; ++t)
      for (int p = plan.off_ptr[t]; p < plan.off_ptr[t + 1]; ++p) dep[blk_of_pos[plan.off_row[p]]].push_back(blk_of_pos[t]);
    for (int b = nblocks - 1; b >= 0; --b) {
      int lv = 0;
      for (int c : dep[b]) lv = std::max(lv, level[c] + 1);
      level[b] = lv;
      maxlevel = std::max(maxlevel, lv);
    }
  }
  plan.level_ptr.assign(maxlevel + 2, 0);
  for (int b = 0; b < nblocks; ++b) ++plan.level_ptr[level[b] + 1];
  for (int l = 0; l <= maxlevel; ++l) plan.level_ptr[l + 1] += plan.level_ptr[l];
  plan.level_blocks.assign(nblocks, 0);
  {
    std::vector<int> fill(plan.level_ptr.begin(), plan.level_ptr.end() - 1);
    for (int b = 0; b < nblocks; ++b) plan.level_blocks[fill[level[b]]++] = b;
  }
  // off-diagonal pull lists per level
  {
    const int kPullLong = 4096, kPullChunk = 4096, kPullTiny = 8;
    plan.pull_tiny_ptr.assign(1, 0);
    plan.pull_short_ptr.assign(1, 0); plan.pull_long_ptr.assign(1, 0); plan.pull_chunk_ptr.assign(1, 0);
    plan.pull_long_chunk_ptr.assign(1, 0);
    for (int l = 0; l <= maxlevel; ++l) {
      for (int q = plan.level_ptr[l]; q < plan.level_ptr[l + 1]; ++q) {
        const int b = plan.level_blocks[q];
        for (int r = bptr[b]; r < bptr[b + 1]; ++r) {
          const int cnt = plan.offr_ptr[r + 1] - plan.offr_ptr[r];
          if (cnt == 0) continue;
          if (cnt <= kPullTiny) { plan.pull_tiny_rows.push_back(r); continue; }      // one thread per row
          if (cnt <= kPullLong) { plan.pull_short_rows.push_back(r); continue; }     // one warp per row
          const int slot = (int)plan.pull_long_rows.size();
          plan.pull_long_rows.push_back(r);
          for (int c0 = plan.offr_ptr[r]; c0 < plan.offr_ptr[r + 1]; c0 += kPullChunk) {
            plan.pull_chunk_row_slot.push_back(slot); plan.pull_chunk_begin.push_back(c0);
          }
          plan.pull_long_chunk_ptr.push_back((int)plan.pull_chunk_begin.size());
        }
      }
      plan.pull_tiny_ptr.push_back((int)plan.pull_tiny_rows.size());
      plan.pull_short_ptr.push_back((int)plan.pull_short_rows.size());
      plan.pull_long_ptr.push_back((int)plan.pull_long_rows.size());
      plan.pull_chunk_ptr.push_back((int)plan.pull_chunk_begin.size());
    }
  }
  // ---- large blocks: column levels for the refactor, row form + row levels for the solves ----
  {
    plan.block_big.assign(nblocks, 0);
    for (int b = 0; b < nblocks; ++b) if (bptr[b + 1] - bptr[b] > kBigBlock) { plan.block_big[b] = 1; plan.big_blocks.push_back(b); }
    plan.acol_dst.assign(plan.acol_row.size(), 0);
    plan.Lr_ptr.assign(n + 1, 0); plan.Ur_ptr.assign(n + 1, 0);
    std::vector<int> clev(n, 0);
    int max_clev = -1;
    std::vector<char> dense(n, 0);
    // destination of every A entry inside the factor arrays (all blocks)
    for (int k = 0; k < n; ++k) {
      const int ub = plan.Up[k], ue = plan.Up[k + 1] - 1, lb = plan.Lp[k], le = plan.Lp[k + 1];
      for (int q = plan.acol_ptr[k]; q < plan.acol_ptr[k + 1]; ++q) {
        const int r = plan.acol_row[q];
        if (r == k) plan.acol_dst[q] = ue;
        else if (r < k) plan.acol_dst[q] = (int)(std::lower_bound(plan.Ui.begin() + ub, plan.Ui.begin() + ue, r) - plan.Ui.begin());
        else plan.acol_dst[q] = ~(int)(std::lower_bound(plan.Li.begin() + lb, plan.Li.begin() + le, r) - plan.Li.begin());
      }
    }
    build_batch_groups(plan, level);
    // small blocks whose factor (indices + values + one dense column) fits a per-warp shared-memory slice are
    // "staged": block_big = 2.  Slice = 2 (nb + 1) + nl + nu 16-bit indices and nb + nl + nu doubles.
    plan.staged_bytes = 0;
    for (int b = 0; b < nblocks; ++b) {
      const int k0 = bptr[b], k1 = bptr[b + 1], nb = k1 - k0;
      if (nb < 2 || plan.block_big[b]) continue;
      const int nl = plan.Lp[k1] - plan.Lp[k0], nu = plan.Up[k1] - plan.Up[k0];
      const int bytes = 8 * (nb + nl + nu) + 2 * (2 * (nb + 1) + nl + nu + 4);      // doubles + 16-bit local indices
      if (bytes <= kStagedBytes) { plan.block_big[b] = 2; plan.staged_bytes = std::max(plan.staged_bytes, (bytes + 15) / 16 * 16); }
    }
    for (int b : plan.big_blocks) {
      for (int k = bptr[b]; k < bptr[b + 1]; ++k) {
        const int ub = plan.Up[k], ue = plan.Up[k + 1] - 1, lb = plan.Lp[k], le = plan.Lp[k + 1];
        int lv = 0;
        for (int q = ub; q < ue; ++q) lv = std::max(lv, clev[plan.Ui[q]] + 1);
        clev[k] = lv; max_clev = std::max(max_clev, lv);
        dense[k] = (ue - ub) > kDenseCol;
        for (int q = lb; q < le; ++q) ++plan.Lr_ptr[plan.Li[q] + 1];
        for (int q = ub; q < ue; ++q) ++plan.Ur_ptr[plan.Ui[q] + 1];
      }
    }
    for (int r = 0; r < n; ++r) { plan.Lr_ptr[r + 1] += plan.Lr_ptr[r]; plan.Ur_ptr[r + 1] += plan.Ur_ptr[r]; }
    plan.Lr_col.resize(plan.Lr_ptr[n]); plan.Lr_src.resize(plan.Lr_ptr[n]);
    plan.Ur_col.resize(plan.Ur_ptr[n]); plan.Ur_src.resize(plan.Ur_ptr[n]);
    {
      std::vector<int> lf(plan.Lr_ptr.begin(), plan.Lr_ptr.end() - 1), uf(plan.Ur_ptr.begin(), plan.Ur_ptr.end() - 1);
      for (int b : plan.big_blocks)
        for (int k = bptr[b]; k < bptr[b + 1]; ++k) {       // ascending k => columns ascending inside every row
          for (int q = plan.Lp[k]; q < plan.Lp[k + 1]; ++q) { const int d = lf[plan.Li[q]]++; plan.Lr_col[d] = k; plan.Lr_src[d] = q; }
          for (int q = plan.Up[k]; q < plan.Up[k + 1] - 1; ++q) { const int d = uf[plan.Ui[q]]++; plan.Ur_col[d] = k; plan.Ur_src[d] = q; }
        }
    }
    // refactor schedule
    plan.rf_level_ptr.assign(1, 0); plan.rf_dense_ptr.assign(1, 0);
    if (max_clev >= 0) {
      std::vector<std::vector<int>> by_level(max_clev + 1), dense_by_level(max_clev + 1);
      for (int b : plan.big_blocks)
        for (int k = bptr[b]; k < bptr[b + 1]; ++k) (dense[k] ? dense_by_level : by_level)[clev[k]].push_back(k);
      for (int l = 0; l <= max_clev; ++l) {
        plan.rf_cols.insert(plan.rf_cols.end(), by_level[l].begin(), by_level[l].end());
        plan.rf_level_ptr.push_back((int)plan.rf_cols.size());
        plan.rf_dense_cols.insert(plan.rf_dense_cols.end(), dense_by_level[l].begin(), dense_by_level[l].end());
        plan.rf_dense_ptr.push_back((int)plan.rf_dense_cols.size());
      }
    }
    // solve schedules, one block after the other
    plan.fs_short_ptr.assign(1, 0); plan.fs_long_ptr.assign(1, 0); plan.bs_short_ptr.assign(1, 0); plan.bs_long_ptr.assign(1, 0);
    std::vector<int> lev(n, 0);
    for (int b : plan.big_blocks) {
      const int k0 = bptr[b], k1 = bptr[b + 1];
      int maxl = 0;
      for (int r = k0; r < k1; ++r) {
        int lv = 0;
        for (int q = plan.Lr_ptr[r]; q < plan.Lr_ptr[r + 1]; ++q) lv = std::max(lv, lev[plan.Lr_col[q]] + 1);
        lev[r] = lv; maxl = std::max(maxl, lv);
      }
      plan.big_fs_begin.push_back((int)plan.fs_short_ptr.size() - 1);
      {
        std::vector<std::vector<int>> rows(maxl + 1);
        for (int r = k0; r < k1; ++r) if (plan.Lr_ptr[r + 1] > plan.Lr_ptr[r]) rows[lev[r]].push_back(r);
        for (int l = 1; l <= maxl; ++l) {        // level 0 rows have no L entries
          for (int r : rows[l]) ((plan.Lr_ptr[r + 1] - plan.Lr_ptr[r] > kLongRow) ? plan.fs_long_rows : plan.fs_short_rows).push_back(r);
          plan.fs_short_ptr.push_back((int)plan.fs_short_rows.size()); plan.fs_long_ptr.push_back((int)plan.fs_long_rows.size());
        }
      }
      plan.big_fs_end.push_back((int)plan.fs_short_ptr.size() - 1);
      maxl = 0;
      for (int r = k1 - 1; r >= k0; --r) {
        int lv = 0;
        for (int q = plan.Ur_ptr[r]; q < plan.Ur_ptr[r + 1]; ++q) lv = std::max(lv, lev[plan.Ur_col[q]] + 1);
        lev[r] = lv; maxl = std::max(maxl, lv);
      }
      plan.big_bs_begin.push_back((int)plan.bs_short_ptr.size() - 1);
      {
        std::vector<std::vector<int>> rows(maxl + 1);
        for (int r = k0; r < k1; ++r) rows[lev[r]].push_back(r);     // every row divides by its pivot, also level 0
        for (int l = 0; l <= maxl; ++l) {
          for (int r : rows[l]) ((plan.Ur_ptr[r + 1] - plan.Ur_ptr[r] > kLongRow) ? plan.bs_long_rows : plan.bs_short_rows).push_back(r);
          plan.bs_short_ptr.push_back((int)plan.bs_short_rows.size()); plan.bs_long_ptr.push_back((int)plan.bs_long_rows.size());
        }
      }
      plan.big_bs_end.push_back((int)plan.bs_short_ptr.size() - 1);
    }
  }
  // flop count of one refactorization: sum_k 2 |L_k| |U_k(offdiag)| + divisions
  double fl = 0.0;
  for (int k = 0; k < n; ++k) {
    for (int p = plan.Up[k]; p < plan.Up[k + 1] - 1; ++p) {
      const int i = plan.Ui[p];
      fl += 2.0 * (plan.Lp[i + 1] - plan.Lp[i]);
    }
    fl += plan.Lp[k + 1] - plan.Lp[k];
  }
  plan.refactor_flops = fl;
}

void set_batching(bool on) { g_batching = on; }

int analyze_and_factor(int n, const int *rowptr, const int *colind, const double *vals, double pivot_tol, LuPlan &plan,
                       const int *val_index) {
  plan = LuPlan();
  plan.n = n;
  const int nnz = rowptr[n];
  plan.nnz_a = nnz;
  // CSR -> CSC with a back-pointer to the CSR value index
  std::vector<int> Ap(n + 1, 0), Ai(nnz), Aidx(nnz);
  for (int k = 0; k < nnz; ++k) ++Ap[colind[k] + 1];
  for (int j = 0; j < n; ++j) Ap[j + 1] += Ap[j];
  {
    std::vector<int> fill(Ap.begin(), Ap.end() - 1);
    for (int i = 0; i < n; ++i)
      for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) { const int p = fill[colind[k]]++; Ai[p] = i; Aidx[p] = val_index ? val_index[k] : k; }
  }
  // 1. maximum transversal: column matched to each row; Q0[i] = column placed at diagonal position i
  std::vector<int> col_of_row;
  const int nmatch = max_transversal(n, Ap, Ai, col_of_row);
  if (nmatch < n) { plan.structurally_singular = true; return 1; }
  // permuted pattern C = A * Q0 has c(i, i') = a(i, col_of_row[i']): column i' of C is column col_of_row[i'] of A
  std::vector<int> Cp(n + 1, 0), Ci;
  Ci.reserve(nnz);
  for (int jp = 0; jp < n; ++jp) {
    const int j = col_of_row[jp];
    for (int p = Ap[j]; p < Ap[j + 1]; ++p) Ci.push_back(Ai[p]);
    Cp[jp + 1] = (int)Ci.size();
  }
  // 2. strongly connected components of C (symmetric permutation)
  std::vector<int> sperm, bptr;
  strong_components(n, Cp, Ci, sperm, bptr);
  const int nblocks = (int)bptr.size() - 1;
  // 3. per-block fill-reducing ordering (within-block symmetric permutation)
  std::vector<int> where(n, -1), blk_of(n, -1);
  for (int b = 0; b < nblocks; ++b)
    for (int t = bptr[b]; t < bptr[b + 1]; ++t) blk_of[sperm[t]] = b;
  std::vector<int> final_order(n);   // position -> index in C numbering
  {
    std::vector<int> local(n, -1);
    for (int b = 0; b < nblocks; ++b) {
      const int nb = bptr[b + 1] - bptr[b];
      if (nb <= 2) { for (int t = bptr[b]; t < bptr[b + 1]; ++t) final_order[t] = sperm[t]; continue; }
      for (int t = 0; t < nb; ++t) local[sperm[bptr[b] + t]] = t;
      std::vector<std::vector<int>> adj(nb);
      for (int t = 0; t < nb; ++t) {
        const int jc = sperm[bptr[b] + t];
        for (int p = Cp[jc]; p < Cp[jc + 1]; ++p) {
          const int ic = Ci[p];
          if (blk_of[ic] != b || ic == jc) continue;
          adj[t].push_back(local[ic]); adj[local[ic]].push_back(t);
        }
      }
      std::vector<int> ord;
      min_degree(nb, adj, ord);
      for (int t = 0; t < nb; ++t) final_order[bptr[b] + t] = sperm[bptr[b] + ord[t]];
    }
  }
  for (int t = 0; t < n; ++t) where[final_order[t]] = t;   // C index -> position
  // At this point position t holds row final_order[t] (of A) and column col_of_row[final_order[t]].
  std::vector<int> rowpos(n), colpos(n);
  for (int t = 0; t < n; ++t) { rowpos[final_order[t]] = t; colpos[col_of_row[final_order[t]]] = t; }

  // 4. factor each diagonal block; collect off-diagonal entries
  plan.block_ptr = bptr;
  plan.Lp.assign(1, 0); plan.Up.assign(1, 0);
  plan.row_perm.assign(n, -1);   // position -> row of A   (includes pivoting)
  plan.col_perm.assign(n, -1);   // position -> column of A
  std::vector<int> a_row(n), a_col(n);
  for (int t = 0; t < n; ++t) { a_row[t] = final_order[t]; a_col[t] = col_of_row[final_order[t]]; }
  std::vector<int> pre_rowpos(rowpos);    // pre-pivot row positions
  std::vector<int> new_rowpos(n, -1);     // row of A -> final position
  plan.singular = false;
  for (int b = 0; b < nblocks; ++b) {
    const int k0 = bptr[b], nb = bptr[b + 1] - bptr[b];
    std::vector<std::vector<std::pair<int, double>>> cols(nb);
    for (int t = 0; t < nb; ++t) {
      const int j = a_col[k0 + t];
      for (int p = Ap[j]; p < Ap[j + 1]; ++p) {
        const int rp = pre_rowpos[Ai[p]];
        if (rp >= k0 && rp < k0 + nb) cols[t].push_back(std::make_pair(rp - k0, vals[Aidx[p]]));
      }
    }
    BlockLU f;
    factor_block(nb, cols, pivot_tol, f);
    if (f.singular) plan.singular = true;
    for (int r = 0; r < nb; ++r) new_rowpos[a_row[k0 + r]] = k0 + f.pinv[r];
    for (int k = 0; k < nb; ++k) {
      for (int p = f.Lp[k]; p < f.Lp[k + 1]; ++p) { plan.Li.push_back(k0 + f.Li[p]); plan.Lx.push_back(f.Lx[p]); }
      plan.Lp.push_back((int)plan.Li.size());
      for (int p = f.Up[k]; p < f.Up[k + 1]; ++p) { plan.Ui.push_back(k0 + f.Ui[p]); plan.Ux.push_back(f.Ux[p]); }
      plan.Up.push_back((int)plan.Ui.size());
    }
  }
  for (int r = 0; r < n; ++r) plan.row_perm[new_rowpos[r]] = r;
  for (int t = 0; t < n; ++t) plan.col_perm[t] = a_col[t];
  finish_plan(plan, Ap, Ai, Aidx, vals, new_rowpos);
  return plan.singular ? 2 : 0;
}

// ---- import of an external factorization's symbolic result -------------------------------------------------
// What a KLU-enabled build hands over after klu_analyze + klu_factor (klu_extract: P, Q, R = block boundaries,
// patterns of L and U), so that the GPU refactor / solve run on KLU's own ordering and pivot sequence instead
// of this file's restatement of it.  Conventions (the adaptor converts): position t of the permuted matrix
// holds row row_perm[t] and column col_perm[t] of A; the permuted matrix is upper block triangular with
// diagonal blocks [block_ptr[b], block_ptr[b+1]); L and U are CSC over positions with row indices as global
// positions inside the column's block.  Tolerated on input: an explicit unit diagonal in L (dropped), U rows
// in any order with the pivot anywhere in the column (sorted, pivot moved last).  row_scale (by position, may be null):
// the external solver factored diag(1 / row_scale) P A Q (KLU scale = 1 or 2, klu_extract's Rs).
// Returns 0 ok, 3 malformed input (message in *why).
int import_factorization(int n, const int *rowptr, const int *colind, const int *row_perm, const int *col_perm,
                         int nblocks, const int *block_ptr, const int *Lp, const int *Li, const int *Up, const int *Ui,
                         const double *row_scale, LuPlan &plan, const char **why) {
  static const char *msg = "";
  *why = msg;
  plan = LuPlan();
  plan.n = n;
  const int nnz = rowptr[n];
  plan.nnz_a = nnz;
  if (nblocks < 1 || block_ptr[0] != 0 || block_ptr[nblocks] != n) { *why = "block_ptr must run from 0 to n"; return 3; }
  std::vector<int> new_rowpos(n, -1), colpos(n, -1);
  for (int t = 0; t < n; ++t) {
    if (row_perm[t] < 0 || row_perm[t] >= n || col_perm[t] < 0 || col_perm[t] >= n) { *why = "permutation entry out of range"; return 3; }
    if (new_rowpos[row_perm[t]] != -1 || colpos[col_perm[t]] != -1) { *why = "row_perm / col_perm is not a permutation"; return 3; }
    new_rowpos[row_perm[t]] = t; colpos[col_perm[t]] = t;
  }
  plan.block_ptr.assign(block_ptr, block_ptr + nblocks + 1);
  plan.row_perm.assign(row_perm, row_perm + n);
  plan.col_perm.assign(col_perm, col_perm + n);
  std::vector<int> blk_of_pos(n);
  for (int b = 0; b < nblocks; ++b) {
    if (block_ptr[b + 1] <= block_ptr[b]) { *why = "empty or decreasing block"; return 3; }
    for (int t = block_ptr[b]; t < block_ptr[b + 1]; ++t) blk_of_pos[t] = b;
  }
  plan.Lp.assign(1, 0); plan.Up.assign(1, 0);
  for (int k = 0; k < n; ++k) {
    const int b = blk_of_pos[k], k0 = block_ptr[b], k1 = block_ptr[b + 1];
    std::vector<int> lr, ur;
    for (int q = Lp[k]; q < Lp[k + 1]; ++q) {
      const int r = Li[q];
      if (r == k) continue;                                   // explicit unit diagonal
      if (r < k || r >= k1) { *why = "L entry outside the strictly lower part of its block"; return 3; }
      lr.push_back(r);
    }
    bool diag = false;
    for (int q = Up[k]; q < Up[k + 1]; ++q) {
      const int r = Ui[q];
      if (r == k) { diag = true; continue; }
      if (r > k || r < k0) { *why = "U entry outside the upper part of its block"; return 3; }
      ur.push_back(r);
    }
    if (!diag) { *why = "U column without its pivot"; return 3; }
    std::sort(lr.begin(), lr.end()); std::sort(ur.begin(), ur.end());
    plan.Li.insert(plan.Li.end(), lr.begin(), lr.end()); plan.Lp.push_back((int)plan.Li.size());
    plan.Ui.insert(plan.Ui.end(), ur.begin(), ur.end()); plan.Ui.push_back(k); plan.Up.push_back((int)plan.Ui.size());
  }
  plan.Lx.assign(plan.Li.size(), 0.0); plan.Ux.assign(plan.Ui.size(), 0.0);
  // CSC of A with CSR value indices
  std::vector<int> Ap(n + 1, 0), Ai(nnz), Aidx(nnz);
  for (int k = 0; k < nnz; ++k) ++Ap[colind[k] + 1];
  for (int j = 0; j < n; ++j) Ap[j + 1] += Ap[j];
  {
    std::vector<int> fill(Ap.begin(), Ap.end() - 1);
    for (int i = 0; i < n; ++i)
      for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) { const int q = fill[colind[k]]++; Ai[q] = i; Aidx[q] = k; }
  }
  // every entry of A must fall on the factor pattern of its block or above the block diagonal
  for (int t = 0; t < n; ++t) {
    const int j = plan.col_perm[t], b = blk_of_pos[t];
    for (int q = Ap[j]; q < Ap[j + 1]; ++q) {
      const int rp = new_rowpos[Ai[q]];
      if (blk_of_pos[rp] > b) { *why = "the permuted matrix is not upper block triangular"; return 3; }
      if (blk_of_pos[rp] != b || rp == t) continue;
      const std::vector<int> &ix = rp < t ? plan.Ui : plan.Li;
      const int lo = rp < t ? plan.Up[t] : plan.Lp[t], hi = rp < t ? plan.Up[t + 1] - 1 : plan.Lp[t + 1];
      if (!std::binary_search(ix.begin() + lo, ix.begin() + hi, rp)) { *why = "an entry of A is missing from the L / U pattern"; return 3; }
    }
  }
  if (row_scale) {
    for (int t = 0; t < n; ++t) if (!(row_scale[t] > 0.0)) { *why = "row_scale must be positive"; return 3; }
    plan.row_scale.assign(row_scale, row_scale + n);
    plan.nz_rowpos.resize(nnz);
    for (int i = 0; i < n; ++i) for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) plan.nz_rowpos[k] = new_rowpos[i];
  }
  finish_plan(plan, Ap, Ai, Aidx, nullptr, new_rowpos);
  return 0;
}

// CPU reference of the refactor + solve on a fixed plan (used by the first solve and by tests of the
// plan itself; the product's per-iteration path is the GPU one).
// Host execution of the batched-group programs (what lu_refactor_batched_kernel / lu_solve_batched_kernel do per
// lane) against a plain left-looking refactorization and triangular solves of the same blocks on the same pattern.
// out[0] = groups, out[1] = batched blocks, out[2] = largest deviation relative to the largest factor entry (refactor),
// out[3] = same for the solve.  Used by the CPU tests: the programs are checked without a GPU.
void batch_selfcheck_host(const LuPlan &p, const double *vals, double *out) {
  out[0] = (double)p.batch.size(); out[1] = 0; out[2] = 0; out[3] = 0;
  for (const LuPlan::BatchGroup &g : p.batch) {
    const int nblk = (int)g.blocks.size(), ns = g.nu + g.nl;
    out[1] += nblk;
    for (int j = 0; j < nblk; ++j) {
      const int b = g.blocks[j], k0 = p.block_ptr[b], k1 = p.block_ptr[b + 1], u0 = p.Up[k0], l0 = p.Lp[k0];
      // program
      std::vector<double> v(ns, 0.0);
      for (int e = 0; e < g.na; ++e) v[g.a_dst[e]] = vals[g.a_src[(size_t)e * nblk + j]];
      for (int bi = 0; bi < g.rf_bundles; ++bi) {
        double d[kBundle], x[kBundle], y[kBundle]; int ty[kBundle], ds[kBundle];
        for (int k = 0; k < kBundle; ++k) {
          const unsigned short *o = &g.rf_prog[((size_t)bi * kBundle + k) * 4];
          ty[k] = o[0] >> 14; ds[k] = o[0] & 0x3fff; d[k] = v[ds[k]]; x[k] = v[o[1]]; y[k] = v[o[2]];
        }
        for (int k = 0; k < kBundle; ++k) {
          if (ty[k] == kOpFnma) v[ds[k]] = d[k] - x[k] * y[k];
          else if (ty[k] == kOpDiv) v[ds[k]] = d[k] / x[k];
        }
      }
      // plain left-looking refactorization of the block with a dense work column
      std::vector<double> Lx(g.nl), Ux(g.nu), x(g.nb);
      for (int k = k0; k < k1; ++k) {
        std::fill(x.begin(), x.end(), 0.0);
        for (int q = p.acol_ptr[k]; q < p.acol_ptr[k + 1]; ++q) x[p.acol_row[q] - k0] = vals[p.acol_src[q]];
        const int ue = p.Up[k + 1] - 1;
        for (int q = p.Up[k]; q < ue; ++q) {
          const int i = p.Ui[q]; const double u = x[i - k0];
          Ux[q - u0] = u;
          for (int t = p.Lp[i]; t < p.Lp[i + 1]; ++t) x[p.Li[t] - k0] -= Lx[t - l0] * u;
        }
        Ux[ue - u0] = x[k - k0];
        for (int q = p.Lp[k]; q < p.Lp[k + 1]; ++q) Lx[q - l0] = x[p.Li[q] - k0] / x[k - k0];
      }
      double mx = 0.0, dev = 0.0;
      for (int s2 = 0; s2 < g.nu; ++s2) { mx = std::max(mx, std::fabs(Ux[s2])); dev = std::max(dev, std::fabs(Ux[s2] - v[s2])); }
      for (int s2 = 0; s2 < g.nl; ++s2) { mx = std::max(mx, std::fabs(Lx[s2])); dev = std::max(dev, std::fabs(Lx[s2] - v[g.nu + s2])); }
      out[2] = std::max(out[2], dev / (mx > 0 ? mx : 1.0));
      // solve: program against direct substitution, right-hand side 1, 2, 3, ...
      std::vector<double> y(g.nb), z(g.nb);
      for (int i = 0; i < g.nb; ++i) y[i] = z[i] = 1.0 + i;
      for (int bi = 0; bi < g.sv_bundles; ++bi) {
        double d[kBundle], xx[kBundle], yy[kBundle]; int ty[kBundle], ds[kBundle];
        for (int k = 0; k < kBundle; ++k) {
          const unsigned short *o = &g.sv_prog[((size_t)bi * kBundle + k) * 4];
          ty[k] = o[0] >> 14; ds[k] = o[0] & 0x3fff; d[k] = y[ds[k]]; xx[k] = v[o[1]]; yy[k] = y[o[2]];
        }
        for (int k = 0; k < kBundle; ++k) {
          if (ty[k] == kOpFnma) y[ds[k]] = d[k] - xx[k] * yy[k];
          else if (ty[k] == kOpDiv) y[ds[k]] = d[k] / xx[k];
        }
      }
      for (int k = k0; k < k1; ++k) for (int q = p.Lp[k]; q < p.Lp[k + 1]; ++q) z[p.Li[q] - k0] -= Lx[q - l0] * z[k - k0];
      for (int k = k1 - 1; k >= k0; --k) {
        const int ue = p.Up[k + 1] - 1;
        z[k - k0] /= Ux[ue - u0];
        for (int q = p.Up[k]; q < ue; ++q) z[p.Ui[q] - k0] -= Ux[q - u0] * z[k - k0];
      }
      double zm = 0.0, zd = 0.0;
      for (int i = 0; i < g.nb; ++i) { zm = std::max(zm, std::fabs(z[i])); zd = std::max(zd, std::fabs(z[i] - y[i])); }
      out[3] = std::max(out[3], zd / (zm > 0 ? zm : 1.0));
    }
  }
}

void solve_host(const LuPlan &p, const double *b, double *x) {
  const int n = p.n;
  std::vector<double> y(n);
  for (int t = 0; t < n; ++t) y[t] = b[p.row_perm[t]];
  const int nblocks = (int)p.block_ptr.size() - 1;
  for (int bb = nblocks - 1; bb >= 0; --bb) {
    const int k0 = p.block_ptr[bb], k1 = p.block_ptr[bb + 1];
    for (int k = k0; k < k1; ++k) {
      const double yk = y[k];
      for (int q = p.Lp[k]; q < p.Lp[k + 1]; ++q) y[p.Li[q]] -= p.Lx[q] * yk;
    }
    for (int k = k1 - 1; k >= k0; --k) {
      const int d = p.Up[k + 1] - 1;
      y[k] /= p.Ux[d];
      const double yk = y[k];
      for (int q = p.Up[k]; q < d; ++q) y[p.Ui[q]] -= p.Ux[q] * yk;
    }
    // push this block's solution into the rows of earlier blocks (off-diagonal columns)
    for (int k = k0; k < k1; ++k)
      for (int q = p.off_ptr[k]; q < p.off_ptr[k + 1]; ++q) y[p.off_row[q]] -= p.off_val_host[q] * y[k];
  }
  for (int t = 0; t < n; ++t) x[p.col_perm[t]] = y[t];
}

}  // namespace lu
}  // namespace xb
