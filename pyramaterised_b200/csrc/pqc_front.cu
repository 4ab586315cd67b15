// pqc_front.cu -- the "front" planner: gate application (circuit.py:118-125) scheduled by
// dependency fronts instead of by template layers, for k_tile_pipe (pqc_pipe.cu).
//
// The block planner of pqc_v1.cu closes a pass whenever the next commuting block needs a bit
// that is not in the tile, which costs the brick-wall circuits dearly: XXZ 16q x 16 needs 43
// passes, the CNOT-chain ansatz two per layer, NPQC one sweep per qubit and layer.  Here the op
// list is a DAG (op j waits for every earlier op it does not commute with) and a pass takes
// EVERYTHING inside the light cone of its 12 tile bits:
//   * tile: the 4 low amplitude bits plus 8 more, grown greedily by how many mixing ops the
//     closure of the tile can execute;
//   * sweeps: any 4 tile bits in registers (k_tile_pipe's geometry is data), chosen one after the
//     other as the set that lets the most ready ops run -- e.g. the 4-bond "diamonds" of a
//     brick-wall, or ALL layers of an NPQC qubit at once (its CZ partners never mix);
//   * diagonal ops run wherever they are ready; X / CNOT are index permutations folded into the
//     load or the store of a sweep whose registers hold their target;
//   * inside a sweep ops are levelled (same level = mutually commuting) and the one-qubit gates of
//     a level are packed into 4-slot layer ops (R_x | R_y, H | R_z in tangent form); same-angle
//     R_zz of a level become one table-lookup phase.
// Supported op kinds: RX RY RZ H X CNOT CZ RZZ, RYY*RXX pairs of one shared parameter, IDENT.
// Anything else keeps the block planner's plan.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <queue>
#include <tuple>

#include "pqc_common.cuh"

namespace {

struct FOp {
  int kind = 0, b0 = -1, b1 = -1, param = -1;
  double scale = 1.0, offset = 0.0;
  uint32_t support = 0, mix = 0, tgt = 0, ctl = 0;
  bool diag = false, pauli = false, two = false;
  uint32_t px = 0, pz = 0, px2 = 0, pz2 = 0;
  std::vector<int> ids;
};

bool pcomm(uint32_t x1, uint32_t z1, uint32_t x2, uint32_t z2) {
  return ((__builtin_popcount(x1 & z2) + __builtin_popcount(z1 & x2)) & 1) == 0;
}

bool is_perm(const FOp& o) { return o.kind == PQC_OP_X || o.kind == PQC_OP_CNOT; }

bool fcommute(const FOp& a, const FOp& b) {
  if (!(a.support & b.support)) return true;
  if (a.diag && b.diag) return true;
  if (a.pauli && b.pauli) {
    bool ok = pcomm(a.px, a.pz, b.px, b.pz);
    if (a.two) ok = ok && pcomm(a.px2, a.pz2, b.px, b.pz);
    if (b.two) ok = ok && pcomm(a.px, a.pz, b.px2, b.pz2);
    if (a.two && b.two) ok = ok && pcomm(a.px2, a.pz2, b.px2, b.pz2);
    return ok;
  }
  // X / CNOT: diagonal in the control's Z basis, an X on the target
  for (int s = 0; s < 2; ++s) {
    const FOp& p = s ? b : a;
    const FOp& o = s ? a : b;
    if (!is_perm(p)) continue;
    if (is_perm(o)) return !(p.tgt & o.ctl) && !(p.ctl & o.tgt);
    if (o.diag) return !(o.support & p.tgt);
    if (o.pauli && !o.two) return !(o.support & p.ctl) && !(o.pz & p.tgt);
    return false;
  }
  return false;
}

bool make_fop(const pqc_op& op, int n, int idx, FOp& c) {
  c = FOp();
  c.ids.push_back(idx);
  c.kind = op.kind;
  c.b0 = n - 1 - op.q0;
  c.b1 = op.q1 >= 0 ? n - 1 - op.q1 : -1;
  c.param = op.param;
  c.scale = op.scale;
  c.offset = op.offset;
  const uint32_t m0 = 1u << c.b0, m1 = c.b1 >= 0 ? 1u << c.b1 : 0u;
  c.support = m0 | m1;
  switch (op.kind) {
    case PQC_OP_RX: c.mix = m0; c.pauli = true; c.px = m0; break;
    case PQC_OP_RY: c.mix = m0; c.pauli = true; c.px = m0; c.pz = m0; break;
    case PQC_OP_RZ: c.diag = true; c.pauli = true; c.pz = m0; break;
    case PQC_OP_H: c.mix = m0; break;
    case PQC_OP_X: c.tgt = m0; break;
    case PQC_OP_IDENT: c.diag = true; break;
    case PQC_OP_CNOT: c.ctl = m0; c.tgt = m1; break;
    case PQC_OP_CZ: c.diag = true; break;
    case PQC_OP_RXX: c.mix = m0 | m1; c.pauli = true; c.px = m0 | m1; break;
    case PQC_OP_RYY: c.mix = m0 | m1; c.pauli = true; c.px = m0 | m1; c.pz = m0 | m1; break;
    case PQC_OP_RZZ: c.diag = true; c.pauli = true; c.pz = m0 | m1; break;
    default: return false;
  }
  return true;
}

bool same_angle(const FOp& a, const FOp& b) {
  return a.param == b.param && a.scale == b.scale && a.offset == b.offset;
}

uint32_t f_swz(uint32_t i) { return i ^ (((i >> 3) ^ (i >> 6) ^ (i >> 9)) & 7u); }

struct Front {
  pqc_program* prog;
  int n;
  std::vector<FOp> ops;
  std::vector<std::vector<int>> preds;
  std::vector<char> done;
  std::vector<int> stamp;
  std::vector<int> lvl;          // level of an op inside the sweep being simulated (valid with stamp)
  std::vector<double> wgt;       // what executing a mixing op / permutation is worth (see below)
  std::vector<uint64_t> desc;    // transitive successors of every op, N x dw bits
  int dw = 0;
  int cur_stamp = 0;
  // worklist forms of the two simulations below (same results, cost proportional to what a
  // simulation touches instead of to the whole pending list; PQC_FRONT_CHECK=1 runs both forms
  // and aborts on any difference)
  std::vector<std::vector<int>> succ;
  std::vector<int> live;         // predecessors of an op that are not done yet
  std::vector<int> cnt, cstamp;  // of those, the ones not taken in the current simulation (valid
                                 // when cstamp == cur_stamp)
  std::vector<int> cand_mark;    // == cand_id: member of the current pass' closure
  int cand_id = 0;
  std::vector<int> pfront;       // pending ops whose predecessors are all done (per pass)
  std::vector<int> sfront;       // the same among the closure's not-done ops (per sweep choice)
  bool worklist = true;          // false: weights are not all 1.0 (PQC_FRONT_ALPHA), sums would
                                 // depend on the visiting order -- keep the scans
  bool check = false;

  bool ready(int i) const {
    for (int p : preds[i])
      if (!done[p] && stamp[p] != cur_stamp) return false;
    return true;
  }

  void mark_done(int i) {
    if (done[i]) return;
    done[i] = 1;
    for (int j : succ[i]) --live[j];
  }

  // ops of `from` that are not done and whose predecessors are all done, in program order
  void frontier_of(const std::vector<int>& from, std::vector<int>& out) const {
    out.clear();
    for (int i : from)
      if (!done[i] && live[i] == 0) out.push_back(i);
  }

  // op i was taken by the current simulation: true when that was j's last missing predecessor
  bool last_pred_taken(int j) {
    if (cstamp[j] != cur_stamp) {
      cstamp[j] = cur_stamp;
      cnt[j] = live[j];
    }
    return --cnt[j] == 0;
  }

  // closure_weight_scan as a worklist from `pfront`: the closure is the least fixed point of
  // "fits the tile and every predecessor is done or inside", whatever the visiting order; the
  // weights are all 1.0 here, so the sum does not depend on the order either
  double closure_weight_fast(uint32_t tile, std::vector<int>* out) {
    ++cur_stamp;
    double w = 0;
    std::vector<int> work;
    for (int i : pfront)
      if (!((ops[i].mix | ops[i].tgt) & ~tile)) work.push_back(i);
    while (!work.empty()) {
      const int i = work.back();
      work.pop_back();
      const FOp& o = ops[i];
      stamp[i] = cur_stamp;
      if (o.mix || o.tgt) w += wgt[i];
      if (out) out->push_back(i);
      for (int j : succ[i])
        if (last_pred_taken(j) && !((ops[j].mix | ops[j].tgt) & ~tile)) work.push_back(j);
    }
    return w;
  }

  double closure_weight(uint32_t tile, const std::vector<int>& pending, std::vector<int>* out) {
    if (!worklist) return closure_weight_scan(tile, pending, out);
    if (!check) return closure_weight_fast(tile, out);
    std::vector<int> a, b;
    const double wa = closure_weight_scan(tile, pending, &a);
    const double wb = closure_weight_fast(tile, &b);
    std::sort(a.begin(), a.end());
    std::sort(b.begin(), b.end());
    if (wa != wb || a != b) {
      fprintf(stderr, "front planner: closure worklist differs from the scan (%g / %g, %zu / %zu ops)\n",
              wa, wb, a.size(), b.size());
      abort();
    }
    if (out) out->insert(out->end(), b.begin(), b.end());
    return wb;
  }

  // mixing ops + permutations the closure of `tile` can execute (any number of sweeps)
  double closure_weight_scan(uint32_t tile, const std::vector<int>& pending, std::vector<int>* out) {
    ++cur_stamp;
    double w = 0;
    bool changed = true;
    std::vector<int> rest = pending, next;
    while (changed) {
      changed = false;
      next.clear();
      for (int i : rest) {
        const FOp& o = ops[i];
        const bool fits = !((o.mix | o.tgt) & ~tile);
        if (fits && ready(i)) {
          stamp[i] = cur_stamp;
          if (o.mix || o.tgt) w += wgt[i];
          if (out) out->push_back(i);
          changed = true;
        } else {
          next.push_back(i);
        }
      }
      rest.swap(next);
    }
    return w;
  }

  struct SweepSim {
    std::vector<int> pre, body, post;
    double weight = 0;
    int ntrig = 0, ntables = 0;    // trig entries / linear-form tables the emission will need
  };

  // what one sweep with register bits R can execute: permutations first (folded into the load),
  // then every ready op whose mixing bits are register bits, then permutations (folded into the
  // store).  `cand` = the pass' closure, in program order.
  void sim_sweep_scan(uint32_t R, const std::vector<int>& cand, int op_room, int trig_room,
                      int wt_room, SweepSim& S) {
    ++cur_stamp;
    S.pre.clear(); S.body.clear(); S.post.clear();
    S.weight = 0;
    int nops = 0, ntrig = 0;
    // R_zz ops become one table-lookup phase per (angle, level): count exactly what the emission
    // will create
    std::vector<std::pair<int, int>> zzkeys;
    auto level_of = [&](int i) {
      int lv = 1;
      for (int p : preds[i])
        if (!done[p] && stamp[p] == cur_stamp && lvl[p] > 0) lv = std::max(lv, lvl[p] + 1);
      return lv;
    };
    auto trig_need = [&](const FOp& o) {
      if (o.kind == PQC_OP_RZZ) return 2;
      if (o.kind == PQC_OP_H || o.kind == PQC_OP_CZ || o.kind == PQC_OP_IDENT || is_perm(o)) return 0;
      return 1;
    };
    auto room = [&](int i, const FOp& o) {
      if (nops + 1 > op_room) return false;
      if (ntrig + trig_need(o) > trig_room) return false;
      if (o.kind == PQC_OP_RZZ) {
        const std::pair<int, int> key(o.param, level_of(i));
        if (std::find(zzkeys.begin(), zzkeys.end(), key) == zzkeys.end() &&
            (int)zzkeys.size() + 1 > wt_room)
          return false;
      }
      return true;
    };
    auto take = [&](int i, std::vector<int>& where, bool body) {
      const FOp& o = ops[i];
      lvl[i] = body ? level_of(i) : 0;
      stamp[i] = cur_stamp;
      where.push_back(i);
      ++nops;
      ntrig += trig_need(o);
      if (o.kind == PQC_OP_RZZ) {
        const std::pair<int, int> key(o.param, lvl[i]);
        if (std::find(zzkeys.begin(), zzkeys.end(), key) == zzkeys.end()) zzkeys.push_back(key);
      }
      if (o.mix || o.tgt) S.weight += wgt[i];
    };
    std::vector<int> rest = cand, next;
    for (int phase = 0; phase < 3; ++phase) {
      bool changed = true;
      while (changed) {
        changed = false;
        next.clear();
        for (int i : rest) {
          const FOp& o = ops[i];
          if (done[i] || stamp[i] == cur_stamp) continue;
          bool ok;
          if (is_perm(o)) ok = phase != 1 && !(o.tgt & ~R);
          else ok = phase == 1 && !(o.mix & ~R);
          if (ok && ready(i) && room(i, o)) {
            take(i, phase == 0 ? S.pre : (phase == 1 ? S.body : S.post), phase == 1);
            changed = true;
          } else {
            next.push_back(i);
          }
        }
        rest.swap(next);
      }
    }
    S.ntrig = ntrig;
    S.ntables = (int)zzkeys.size();
  }

  // sim_sweep_scan from the ready frontier (`sfront`) instead of over the whole closure.  The scan
  // visits the closure in program order and every predecessor of an op precedes it, so one round
  // of a phase takes whatever the phase can take, in ascending op index; a min-heap over the
  // ready ops pops in exactly that order (an op enters when its last predecessor is taken, and
  // that predecessor has a smaller index), refusals (register set, room) are final within a
  // phase, and the ready ops a phase refused are the next phase's start.  Same takes, same order,
  // same room / level / table bookkeeping.
  void sim_sweep_fast(uint32_t R, int op_room, int trig_room, int wt_room, SweepSim& S) {
    ++cur_stamp;
    S.pre.clear(); S.body.clear(); S.post.clear();
    S.weight = 0;
    int nops = 0, ntrig = 0;
    std::vector<std::pair<int, int>> zzkeys;
    auto level_of = [&](int i) {
      int lv = 1;
      for (int p : preds[i])
        if (!done[p] && stamp[p] == cur_stamp && lvl[p] > 0) lv = std::max(lv, lvl[p] + 1);
      return lv;
    };
    auto trig_need = [&](const FOp& o) {
      if (o.kind == PQC_OP_RZZ) return 2;
      if (o.kind == PQC_OP_H || o.kind == PQC_OP_CZ || o.kind == PQC_OP_IDENT || is_perm(o)) return 0;
      return 1;
    };
    auto room = [&](int i, const FOp& o) {
      if (nops + 1 > op_room) return false;
      if (ntrig + trig_need(o) > trig_room) return false;
      if (o.kind == PQC_OP_RZZ) {
        const std::pair<int, int> key(o.param, level_of(i));
        if (std::find(zzkeys.begin(), zzkeys.end(), key) == zzkeys.end() &&
            (int)zzkeys.size() + 1 > wt_room)
          return false;
      }
      return true;
    };
    std::priority_queue<int, std::vector<int>, std::greater<int>> heap;
    std::vector<int> refused;
    for (int i : sfront) heap.push(i);
    for (int phase = 0; phase < 3; ++phase) {
      std::vector<int>& where = phase == 0 ? S.pre : (phase == 1 ? S.body : S.post);
      refused.clear();
      while (!heap.empty()) {
        const int i = heap.top();
        heap.pop();
        const FOp& o = ops[i];
        bool ok;
        if (is_perm(o)) ok = phase != 1 && !(o.tgt & ~R);
        else ok = phase == 1 && !(o.mix & ~R);
        if (!(ok && room(i, o))) {
          refused.push_back(i);
          continue;
        }
        lvl[i] = phase == 1 ? level_of(i) : 0;
        stamp[i] = cur_stamp;
        where.push_back(i);
        ++nops;
        ntrig += trig_need(o);
        if (o.kind == PQC_OP_RZZ) {
          const std::pair<int, int> key(o.param, lvl[i]);
          if (std::find(zzkeys.begin(), zzkeys.end(), key) == zzkeys.end()) zzkeys.push_back(key);
        }
        if (o.mix || o.tgt) S.weight += wgt[i];
        for (int j : succ[i])
          if (last_pred_taken(j) && cand_mark[j] == cand_id) heap.push(j);
      }
      for (int i : refused) heap.push(i);
    }
    S.ntrig = ntrig;
    S.ntables = (int)zzkeys.size();
  }

  void sim_sweep(uint32_t R, const std::vector<int>& cand, int op_room, int trig_room, int wt_room,
                 SweepSim& S) {
    if (!worklist) return sim_sweep_scan(R, cand, op_room, trig_room, wt_room, S);
    if (!check) return sim_sweep_fast(R, op_room, trig_room, wt_room, S);
    SweepSim T;
    sim_sweep_scan(R, cand, op_room, trig_room, wt_room, T);
    sim_sweep_fast(R, op_room, trig_room, wt_room, S);
    if (T.pre != S.pre || T.body != S.body || T.post != S.post || T.weight != S.weight ||
        T.ntrig != S.ntrig || T.ntables != S.ntables) {
      fprintf(stderr, "front planner: sweep worklist differs from the scan (R = %x: %zu+%zu+%zu / "
              "%zu+%zu+%zu ops)\n", R, T.pre.size(), T.body.size(), T.post.size(), S.pre.size(),
              S.body.size(), S.post.size());
      abort();
    }
  }
};

struct TrigKey {
  int cls, param;
  double scale, offset;
  bool operator<(const TrigKey& o) const {
    return std::tie(cls, param, scale, offset) < std::tie(o.cls, o.param, o.scale, o.offset);
  }
};

}  // namespace

void pqc_pipe_fill_load_tables(PipePlan& pp) {
  // thread bits 0-3 on the tile positions of amplitude bits 0-3 (256-byte runs per half warp),
  // thread bits 4-7 and the copy index j on the other positions in ascending order
  int tpos[12], nt = 0;
  for (int b = 0; b < 4; ++b)
    for (int p = 0; p < 12; ++p)
      if (pp.lbit[p] == b) tpos[nt++] = p;
  for (int p = 0; p < 12; ++p)
    if (pp.lbit[p] >= 4) tpos[nt++] = p;
  for (int h = 0; h < 2; ++h)
    for (int v = 0; v < 16; ++v) {
      uint32_t idx = 0, amp = 0;
      for (int i = 0; i < 4; ++i)
        if ((v >> i) & 1) {
          idx |= 1u << tpos[4 * h + i];
          amp |= 1u << pp.lbit[tpos[4 * h + i]];
        }
      pp.ld_slot[h][v] = (uint16_t)(f_swz(idx) << 4);
      pp.ld_amp[h][v] = amp;
    }
  for (int k = 0; k < 4; ++k) {
    pp.ld_sr[k] = (uint16_t)(f_swz(1u << tpos[8 + k]) << 4);
    pp.ld_r[k] = 1u << pp.lbit[tpos[8 + k]];
  }
}

// Plans PQC.run for k_tile_pipe.  Appends its passes to prog->v1_passes / prog->h_pipe and its
// trig jobs to `tjobs`; fills prog->front_run.  Returns false (and leaves everything untouched)
// when the circuit has an op kind it does not handle.
#define FRONT_FAIL(code)                                                        \
  do {                                                                          \
    if (getenv("PQC_FRONT_DEBUG")) fprintf(stderr, "front planner: fail %d\n", code); \
    return false;                                                               \
  } while (0)
static bool front_impl(pqc_program* prog, std::vector<TrigJob>& tjobs) {
  const int n = prog->n;
  if (n < 12) FRONT_FAIL(1);
  Front F;
  F.prog = prog;
  F.n = n;
  // ---- ops, with R_yy * R_xx of one shared parameter on one pair fused into an XY rotation
  {
    std::vector<FOp> run;
    int cur = -2;
    bool bad = false;
    auto flush = [&]() {
      std::vector<bool> dead(run.size(), false);
      bool all_comm = run.size() >= 2;
      for (size_t i = 0; i < run.size() && all_comm; ++i) {
        if (!run[i].pauli) all_comm = false;
        for (size_t j = i + 1; j < run.size() && all_comm; ++j)
          if (!fcommute(run[i], run[j])) all_comm = false;
      }
      if (all_comm)
        for (size_t i = 0; i < run.size(); ++i) {
          if (dead[i] || run[i].kind != PQC_OP_RYY) continue;
          for (size_t j = 0; j < run.size(); ++j) {
            if (dead[j] || run[j].kind != PQC_OP_RXX || !same_angle(run[i], run[j])) continue;
            if (run[i].support != run[j].support) continue;
            run[i].kind = PQC_K_RXY;
            run[i].two = true;
            run[i].px2 = run[j].px;
            run[i].pz2 = run[j].pz;
            run[i].ids.insert(run[i].ids.end(), run[j].ids.begin(), run[j].ids.end());
            dead[j] = true;
            break;
          }
        }
      for (size_t i = 0; i < run.size(); ++i) {
        if (dead[i]) continue;
        if (run[i].kind == PQC_OP_RXX || run[i].kind == PQC_OP_RYY) bad = true;   // unpaired
        F.ops.push_back(run[i]);
      }
      run.clear();
    };
    for (size_t oi = 0; oi < prog->ops.size(); ++oi) {
      const pqc_op& op = prog->ops[oi];
      if (op.group != cur) {
        flush();
        cur = op.group;
      }
      FOp c;
      if (!make_fop(op, n, (int)oi, c)) FRONT_FAIL(2);
      run.push_back(c);
    }
    flush();
    if (bad) FRONT_FAIL(3);
  }
  const int N = (int)F.ops.size();
  if (N == 0) FRONT_FAIL(4);
  // ---- DAG: op j waits for every earlier op it shares a bit with and does not commute with.
  // Per bit only the ops since the last "wall" (an op nothing later on that bit commutes past is
  // not tracked; the plain scan is quadratic in the ops per bit, which is small).
  F.preds.assign(N, {});
  {
    std::vector<std::vector<int>> on_bit(n);
    for (int j = 0; j < N; ++j) {
      std::vector<int>& pj = F.preds[j];
      for (int b = 0; b < n; ++b) {
        if (!((F.ops[j].support >> b) & 1u)) continue;
        for (int i : on_bit[b])
          if (!fcommute(F.ops[i], F.ops[j])) pj.push_back(i);
      }
      std::sort(pj.begin(), pj.end());
      pj.erase(std::unique(pj.begin(), pj.end()), pj.end());
      for (int b = 0; b < n; ++b)
        if ((F.ops[j].support >> b) & 1u) on_bit[b].push_back(j);
    }
  }
  F.done.assign(N, 0);
  F.stamp.assign(N, 0);
  F.lvl.assign(N, 0);
  // ---- what an op is worth to a pass: itself plus a share of everything that waits for it.  A
  // rotation that many later ops depend on (the first-layer rotations of NPQC's odd qubits, which
  // every CZ partner's chain waits for) should run early even if its own sweep is light: the passes
  // behind it can then run whole chains with full slots.  Descendants by bitsets, reverse order.
  {
    static const double alpha = getenv("PQC_FRONT_ALPHA") ? atof(getenv("PQC_FRONT_ALPHA")) : 0.0;
    const int W = (N + 63) / 64;
    F.dw = W;
    F.desc.assign((size_t)N * W, 0);
    std::vector<uint64_t>& desc = F.desc;
    F.succ.assign(N, {});
    std::vector<std::vector<int>>& succ = F.succ;
    for (int j = 0; j < N; ++j)
      for (int i : F.preds[j]) succ[i].push_back(j);
    F.cand_mark.assign(N, 0);
    F.live.assign(N, 0);
    for (int j = 0; j < N; ++j) F.live[j] = (int)F.preds[j].size();
    F.cnt.assign(N, 0);
    F.cstamp.assign(N, 0);
    F.worklist = alpha == 0.0;
    F.check = getenv("PQC_FRONT_CHECK") != nullptr;
    F.wgt.assign(N, 1.0);
    for (int i = N - 1; i >= 0; --i) {
      uint64_t* di = &desc[(size_t)i * W];
      for (int sidx : succ[i]) {
        const uint64_t* ds = &desc[(size_t)sidx * W];
        for (int w = 0; w < W; ++w) di[w] |= ds[w];
        di[sidx >> 6] |= 1ull << (sidx & 63);
      }
      long cnt = 0;
      for (int w = 0; w < W; ++w) cnt += __builtin_popcountll(di[w]);
      F.wgt[i] = 1.0 + alpha * (double)cnt;
    }
  }

  const size_t tj0 = tjobs.size();
  const uint32_t low = 0xfu;
  int plan_slots = 0;
  std::vector<int> run_list;
  std::vector<int> pending;
  for (int i = 0; i < N; ++i) pending.push_back(i);
  int guard = 0;
  while (!pending.empty()) {
    if (++guard > 4096) FRONT_FAIL(5);
    // ---- tile: grow by closure gain per added bit (candidates: the missing bits of pending ops)
    uint32_t tile = low;
    F.frontier_of(pending, F.pfront);
    // cheap unblockers first: a bit with at most two pending rotations that many ops on OTHER bits
    // wait for (NPQC's odd qubits: one first-layer rotation each, then only CZ partners).  With
    // them in the tile early the later passes run whole chains with every register slot busy and
    // every CZ partner a thread-constant bit.
    uint32_t cheap = 0;                         // every cheap unblocker, in the tile or not
    {
      static const bool seed_on = !getenv("PQC_FRONT_NOSEED");
      std::vector<int> own(n, 0);
      for (int i : pending) {
        const uint32_t m = F.ops[i].mix | F.ops[i].tgt;
        for (int b = 0; b < n; ++b) own[b] += (m >> b) & 1u;
      }
      std::vector<std::pair<double, int>> seeds;
      std::vector<uint64_t> acc(F.dw);
      for (int b = 0; b < n && seed_on; ++b) {
        if (own[b] == 0 || own[b] > 2) continue;
        std::fill(acc.begin(), acc.end(), 0);
        for (int i : pending)
          if (((F.ops[i].mix | F.ops[i].tgt) >> b) & 1u)
            for (int w = 0; w < F.dw; ++w) acc[w] |= F.desc[(size_t)i * F.dw + w];
        long blocked = 0;
        for (int j : pending) {
          const uint32_t m = F.ops[j].mix | F.ops[j].tgt;
          if (m && !((m >> b) & 1u) && ((acc[j >> 6] >> (j & 63)) & 1ull)) ++blocked;
        }
        if (blocked >= 8 * own[b]) {
          cheap |= 1u << b;
          if (!((tile >> b) & 1u)) seeds.push_back(std::make_pair(-(double)blocked / own[b], b));
        }
      }
      std::sort(seeds.begin(), seeds.end());
      for (const auto& sd : seeds)
        if (__builtin_popcount(tile) < 12) tile |= 1u << sd.second;
    }
    while (__builtin_popcount(tile) < 12) {
      const double base = F.closure_weight(tile, pending, nullptr);
      std::vector<uint32_t> cands;
      for (int i : pending) {
        const uint32_t need = (F.ops[i].mix | F.ops[i].tgt) & ~tile;
        if (!need || __builtin_popcount(need) + __builtin_popcount(tile) > 12) continue;
        if (std::find(cands.begin(), cands.end(), need) == cands.end()) cands.push_back(need);
        if (cands.size() >= 96) break;
      }
      if (cands.empty()) break;
      double best = -1.0;
      uint32_t bm = 0;
      for (uint32_t m : cands) {
        const double g = F.closure_weight(tile | m, pending, nullptr) - base;
        const double score = g / __builtin_popcount(m);
        if (score > best) { best = score; bm = m; }
      }
      tile |= bm;
    }
    for (int b = 0; b < n && __builtin_popcount(tile) < 12; ++b) tile |= 1u << b;
    std::vector<int> cand;
    F.closure_weight(tile, pending, &cand);
    std::sort(cand.begin(), cand.end());
    ++F.cand_id;
    for (int i : cand) F.cand_mark[i] = F.cand_id;
    int tb[12], lpos[32], nt = 0;
    for (int b = 0; b < 32; ++b) lpos[b] = -1;
    for (int b = 0; b < n; ++b)
      if ((tile >> b) & 1u) { lpos[b] = nt; tb[nt++] = b; }
    // bits worth holding in registers
    uint32_t useful = 0;
    for (int i : cand) useful |= F.ops[i].mix | F.ops[i].tgt;
    // ---- sweeps
    struct SweepOut { uint32_t R; Front::SweepSim sim; };
    std::vector<SweepOut> sweeps;
    int ops_used = 0, trig_used = 0, wt_used = 0;
    while ((int)sweeps.size() < TP_MAX_SWEEPS - 1) {
      Front::SweepSim best, cur;
      uint32_t bestR = 0;
      best.weight = -1;
      int best_total = -1;
      // candidate register sets: 4 useful tile bits (padded with others when fewer are useful)
      std::vector<int> ub, other;
      for (int p = 0; p < 12; ++p) ((useful >> tb[p]) & 1u ? ub : other).push_back(tb[p]);
      std::vector<uint32_t> Rs;
      {
        // cheap unblockers of the tile that still have a ready rotation: a light sweep of their own
        // (padded with bits that have nothing to do), so that the chains they unblock meet their CZ
        // partners as thread-constant bits later instead of sharing the registers with them
        uint32_t cu = 0;
        for (int i : cand) {
          const FOp& o = F.ops[i];
          if (F.done[i] || !(o.mix | o.tgt) || ((o.mix | o.tgt) & ~cheap)) continue;
          bool rdy = true;
          for (int pr : F.preds[i]) rdy = rdy && F.done[pr];
          if (rdy) cu |= o.mix | o.tgt;
        }
        cu &= tile;
        if (cu) {
          uint32_t R = 0;
          for (int b = 0; b < n && __builtin_popcount(R) < 4; ++b)
            if ((cu >> b) & 1u) R |= 1u << b;
          for (int i = (int)other.size() - 1; i >= 0 && __builtin_popcount(R) < 4; --i) R |= 1u << other[i];
          for (int b : ub)
            if (__builtin_popcount(R) < 4 && !((R >> b) & 1u) && ((cheap >> b) & 1u)) R |= 1u << b;
          for (int b : ub)
            if (__builtin_popcount(R) < 4 && !((R >> b) & 1u)) R |= 1u << b;
          if (__builtin_popcount(R) == 4) Rs.push_back(R);
        }
      }
      if (!Rs.empty()) {
      } else if ((int)ub.size() <= 4) {
        uint32_t R = 0;
        for (int b : ub) R |= 1u << b;
        for (int i = (int)other.size() - 1; i >= 0 && __builtin_popcount(R) < 4; --i) R |= 1u << other[i];
        Rs.push_back(R);
      } else {
        const int m = (int)ub.size();
        for (int a = 0; a < m; ++a)
          for (int b = a + 1; b < m; ++b)
            for (int c = b + 1; c < m; ++c)
              for (int d = c + 1; d < m; ++d)
                Rs.push_back((1u << ub[a]) | (1u << ub[b]) | (1u << ub[c]) | (1u << ub[d]));
      }
      F.frontier_of(cand, F.sfront);
      for (uint32_t R : Rs) {
        F.sim_sweep(R, cand, TP_MAX_OPS - 4 - ops_used, TP_MAX_TRIG - 2 - trig_used,
                    TP_MAX_WT - wt_used, cur);
        const int total = (int)(cur.pre.size() + cur.body.size() + cur.post.size());
        if (cur.weight > best.weight || (cur.weight == best.weight && total > best_total)) {
          best = cur;
          bestR = R;
          best_total = total;
        }
      }
      if (best_total <= 0) break;
      if (best.weight <= 0 && !sweeps.empty()) {
        // only diagonal ops are left in reach: they ride along in the last sweep of the pass
        // (its geometry is irrelevant to them) instead of opening one of their own
        break;
      }
      for (int i : best.pre) F.mark_done(i);
      for (int i : best.body) F.mark_done(i);
      for (int i : best.post) F.mark_done(i);
      const int nz = best.ntables, ntr = best.ntrig;
      ops_used += best_total;
      trig_used += ntr;
      wt_used += nz;
      sweeps.push_back(SweepOut{bestR, best});
      if (ops_used >= TP_MAX_OPS - 8 || trig_used >= TP_MAX_TRIG - 4 || wt_used >= TP_MAX_WT) break;
    }
    if (sweeps.empty()) FRONT_FAIL(6);
    // leftover ready diagonal ops of the closure join the last sweep's body
    {
      bool changed = true;
      ++F.cur_stamp;
      while (changed) {
        changed = false;
        for (int i : cand) {
          const FOp& o = F.ops[i];
          if (F.done[i] || o.mix || o.tgt || !sweeps.back().sim.post.empty()) continue;
          if (ops_used >= TP_MAX_OPS - 4 || trig_used >= TP_MAX_TRIG - 2) continue;
          if (o.kind == PQC_OP_RZZ && wt_used >= TP_MAX_WT) continue;
          bool rdy = true;
          for (int p : F.preds[i]) if (!F.done[p]) rdy = false;
          if (!rdy) continue;
          F.mark_done(i);
          sweeps.back().sim.body.push_back(i);
          ++ops_used;
          trig_used += o.kind == PQC_OP_RZZ ? 2 : 1;
          if (o.kind == PQC_OP_RZZ) ++wt_used;
          changed = true;
        }
      }
    }
    {
      std::vector<int> np;
      for (int i : pending) if (!F.done[i]) np.push_back(i);
      if (np.size() == pending.size()) FRONT_FAIL(7);
      pending.swap(np);
    }
    // the last sweep stores to global memory from registers: lanes must cover amplitude bits
    // 0-3, so a last sweep holding one of them gets an op-less transposition sweep behind it
    if (sweeps.back().R & low) {
      SweepOut d;
      d.R = 0;
      for (int p = 11; p >= 0 && __builtin_popcount(d.R) < 4; --p) d.R |= 1u << tb[p];
      sweeps.push_back(d);
    }
    // ---- emit the pass
    PipePlan pp;
    memset(&pp, 0, sizeof(pp));
    V1Pass ps;
    memset(ps.lbit, 0, sizeof(ps.lbit));
    memset(ps.obit, 0, sizeof(ps.obit));
    ps.tb = 12;
    for (int p = 0; p < 12; ++p) pp.lbit[p] = ps.lbit[p] = tb[p];
    {
      int o = 0;
      for (int b = 0; b < n; ++b)
        if (lpos[b] < 0) ps.obit[o++] = b;
    }
    ps.low_run = 0;
    while (ps.low_run < 12 && ps.lbit[ps.low_run] == ps.low_run) ps.low_run++;
    ps.direct_ok = true;
    ps.front = true;
    pp.nsw = (int)sweeps.size();
    pqc_pipe_fill_load_tables(pp);
    std::map<TrigKey, int> tslot;
    std::vector<TrigJob> tj;
    int ntrig = 0, nwt = 0, nops = 0;
    auto trig_of = [&](int cls, const FOp& o) -> int {
      // cls 0: (tan, cos) of the half angle; 1: (cos, sin) of the half angle; 2: (cos, sin) of
      // the full angle (XY rotation)
      const TrigKey key{cls, o.param, o.scale, o.offset};
      auto it = tslot.find(key);
      if (it != tslot.end()) return it->second;
      TrigJob j;
      memset(&j, 0, sizeof(j));
      j.kind = cls == 2 ? PQC_K_RXY : (cls == 0 ? PQC_OP_RX : PQC_OP_RZ);
      j.param = o.param;
      j.param2 = -1;
      j.slot = ntrig;
      j.pad = cls == 0 ? 1 : 0;
      j.scale = o.scale;
      j.offset = o.offset;
      tj.push_back(j);
      tslot[key] = ntrig;
      return ntrig++;
    };
    // ry(alpha) CZ ry(theta) merged into one rotation: entries (tan, cos) of (theta + alpha) / 2
    // and (theta - alpha) / 2 in two consecutive slots (partner bit 0 / 1)
    auto trig_pair = [&](const FOp& o, double alpha) -> int {
      const int base = ntrig;
      for (int e = 0; e < 2; ++e) {
        TrigJob j;
        memset(&j, 0, sizeof(j));
        j.kind = PQC_OP_RX;
        j.param = o.param;
        j.param2 = -1;
        j.slot = ntrig++;
        j.pad = 1;
        j.scale = o.scale;
        j.offset = o.offset + (e == 0 ? alpha : -alpha);
        tj.push_back(j);
      }
      return base;
    };
    for (int s = 0; s < pp.nsw; ++s) {
      TPSweep& sw = pp.sw[s];
      const uint32_t R = sweeps[s].R;
      const bool last = s + 1 == pp.nsw;
      int rpos[4], tp[8], nr = 0, ntp = 0, kof[32];
      for (int b = 0; b < 32; ++b) kof[b] = -1;
      for (int p = 0; p < 12; ++p)
        if ((R >> tb[p]) & 1u) { kof[tb[p]] = nr; rpos[nr++] = p; }
      if (nr != 4) FRONT_FAIL(8);
      {
        std::vector<int> rest;
        for (int p = 0; p < 12; ++p)
          if (!((R >> tb[p]) & 1u)) rest.push_back(p);
        std::vector<int> order;
        if (last) {
          order = rest;                       // ascending: positions 0-3 (amplitude bits 0-3) first
        } else {
          // lanes 0-2: three positions with distinct residues mod 3 (conflict-free 16-byte slots)
          bool used[12] = {false};
          for (int res = 0; res < 3; ++res)
            for (int p : rest)
              if (!used[p] && p % 3 == res) { order.push_back(p); used[p] = true; break; }
          for (int p : rest)
            if (!used[p]) order.push_back(p);
        }
        for (int p : order) tp[ntp++] = p;
      }
      for (int k = 0; k < 4; ++k) {
        sw.rpos[k] = (uint8_t)rpos[k];
        sw.rs[k] = (uint16_t)(f_swz(1u << rpos[k]) << 4);
      }
      for (int h = 0; h < 2; ++h)
        for (int v = 0; v < 16; ++v) {
          uint32_t idx = 0;
          for (int i = 0; i < 4; ++i)
            if ((v >> i) & 1) idx |= 1u << tp[4 * h + i];
          sw.tt[h][v] = (f_swz(idx) << 4) | (idx << 16);
          if (last) {
            uint32_t amp = 0;
            for (int p = 0; p < 12; ++p)
              if ((idx >> p) & 1u) amp |= 1u << tb[p];
            pp.st_t[h][v] = amp;
          }
        }
      if (last)
        for (int k = 0; k < 4; ++k) pp.st_r[k] = 1u << tb[rpos[k]];
      sw.op_begin = (uint16_t)nops;
      auto push = [&](const TPOp& o) -> bool {
        if (nops >= TP_MAX_OPS) FRONT_FAIL(9);
        pp.ops[nops++] = o;
        return true;
      };
      auto perm_op = [&](const FOp& o) -> TPOp {
        TPOp t;
        memset(&t, 0, sizeof(t));
        t.kind = (uint8_t)o.kind;
        t.a = 0xff;
        if (o.kind == PQC_OP_X) {
          t.b = (uint8_t)kof[o.b0];
        } else {
          t.b = (uint8_t)kof[o.b1];
          if (kof[o.b0] >= 0) t.a = (uint8_t)kof[o.b0];
          t.t[0] = lpos[o.b0] >= 0 ? (uint16_t)lpos[o.b0] : 0xffff;
          t.t[1] = (uint16_t)o.b0;
        }
        return t;
      };
      const Front::SweepSim& sim = sweeps[s].sim;
      for (int i : sim.pre) {
        if (!push(perm_op(F.ops[i]))) FRONT_FAIL(10);
        ps.op_ids.insert(ps.op_ids.end(), F.ops[i].ids.begin(), F.ops[i].ids.end());
      }
      sw.npre = (uint8_t)sim.pre.size();
      // the relabeling of ops [b, e) as the kernel's tp_affine would work it out per thread (reverse:
      // folded into the load, i.e. the inverse order), but once, here: col / v_const by simulation,
      // every CNOT with a thread-fixed control as an injection carried through the later ops
      auto build_aff = [&](int b, int e, bool reverse, TPAff& af, bool global) -> bool {
        uint32_t col[4] = {1u, 2u, 4u, 8u}, vc = 0u, m[TP_MAX_INJ];
        uint8_t src[TP_MAX_INJ];
        int ninj = 0;
        bool ok = true;
        for (int q = 0; q < e - b; ++q) {
          const TPOp& pm = pp.ops[reverse ? e - 1 - q : b + q];
          const int kt = pm.b;
          if (pm.kind == PQC_OP_X) {
            vc ^= 1u << kt;
          } else if (pm.a != 0xff) {
            const int kc = pm.a;
            for (int c = 0; c < 4; ++c) col[c] ^= ((col[c] >> kc) & 1u) << kt;
            vc ^= ((vc >> kc) & 1u) << kt;
            for (int i2 = 0; i2 < ninj; ++i2) m[i2] ^= ((m[i2] >> kc) & 1u) << kt;
          } else {
            const uint8_t sb8 = pm.t[0] != 0xffff ? (uint8_t)pm.t[0] : (uint8_t)(0x80 | pm.t[1]);
            int at = -1;
            for (int i2 = 0; i2 < ninj; ++i2)
              if (src[i2] == sb8) at = i2;
            if (at >= 0) {
              m[at] ^= 1u << kt;
            } else if (ninj < TP_MAX_INJ) {
              src[ninj] = sb8;
              m[ninj++] = 1u << kt;
            } else {
              ok = false;
            }
          }
        }
        memset(&af, 0, sizeof(af));
        if (!ok) return false;
        auto lin16 = [&](uint32_t x) {
          uint32_t r = 0;
          for (int k = 0; k < 4; ++k) if ((x >> k) & 1u) r ^= sw.rs[k];
          return (uint16_t)r;
        };
        auto lin32 = [&](uint32_t x) {
          uint32_t r = 0;
          for (int k = 0; k < 4; ++k) if ((x >> k) & 1u) r ^= 1u << tb[rpos[k]];
          return r;
        };
        for (int c = 0; c < 4; ++c) af.l[c] = lin16(col[c]);
        af.base = lin16(vc);
        af.ninj = (uint8_t)ninj;
        for (int i2 = 0; i2 < ninj; ++i2) {
          af.inj[i2].src = src[i2];
          af.inj[i2].lm = lin16(m[i2]);
        }
        if (global) {
          for (int c = 0; c < 4; ++c) pp.st_q[c] = lin32(col[c]);
          pp.st_base = lin32(vc);
          for (int i2 = 0; i2 < ninj; ++i2) pp.st_gm[i2] = lin32(m[i2]);
        }
        return true;
      };
      if (sw.npre && !build_aff(sw.op_begin, sw.op_begin + sw.npre, true, sw.pre, false)) FRONT_FAIL(21);
      // ---- body: levels of mutually commuting ops
      {
        // peephole: ry(fixed alpha) -> CZ(q, thread-constant bit) -> ry on one register qubit q
        // with nothing else on q in between becomes ONE rotation whose angle depends on the
        // partner bit (the CZ itself turns into the pending Z frame, see PQC_K_CZF)
        std::map<int, std::pair<double, int>> merged;    // ry op -> (alpha, partner amplitude bit)
        std::vector<char> absorbed(N, 0);
        static const bool no_merge = getenv("PQC_FRONT_NOMERGE") != nullptr;
        static const bool no_czf = getenv("PQC_FRONT_NOCZF") != nullptr;
        for (int q = 0; q < n && !no_merge && !no_czf; ++q) {
          if (kof[q] < 0) continue;
          std::vector<int> seq;
          for (int i : sim.body)
            if ((F.ops[i].support >> q) & 1u) seq.push_back(i);
          for (size_t x = 0; x + 2 < seq.size(); ++x) {
            const FOp& A = F.ops[seq[x]];
            const FOp& C = F.ops[seq[x + 1]];
            const FOp& B = F.ops[seq[x + 2]];
            if (A.kind != PQC_OP_RY || A.param >= 0 || absorbed[seq[x]] || merged.count(seq[x])) continue;
            if (C.kind != PQC_OP_CZ || B.kind != PQC_OP_RY) continue;
            const int other = C.b0 == q ? C.b1 : C.b0;
            if (kof[other] >= 0) continue;
            merged[seq[x + 2]] = std::make_pair(A.offset, other);
            absorbed[seq[x]] = absorbed[seq[x + 1]] = 1;
            x += 2;
          }
        }
        auto partner_byte = [&](int bit) -> uint8_t {
          return lpos[bit] >= 0 ? (uint8_t)lpos[bit] : (uint8_t)(0x80 | bit);
        };
        std::map<int, int> level;
        int maxlev = 0;
        for (int i : sim.body) {
          int lv = 1;
          for (int p : F.preds[i]) {
            auto it = level.find(p);
            if (it != level.end()) lv = std::max(lv, it->second + 1);
          }
          level[i] = lv;
          maxlev = std::max(maxlev, lv);
        }
        for (int lv = 1; lv <= maxlev; ++lv) {
          std::vector<int> L;
          for (int i : sim.body)
            if (level[i] == lv && !absorbed[i]) L.push_back(i);
          // 4-slot layer ops: class 0 rx, 1 ry (or the ry / Hadamard mix), 2 rz on a register bit
          bool level_has_h = false;
          for (int i : L) level_has_h = level_has_h || F.ops[i].kind == PQC_OP_H;
          for (int cls = 0; cls < 3; ++cls) {
            std::vector<TPOp> packs;
            for (int i : L) {
              const FOp& o = F.ops[i];
              int code = 0;
              if (cls == 0 && o.kind == PQC_OP_RX) code = 1;
              if (cls == 1 && o.kind == PQC_OP_RY) code = 1;
              if (cls == 1 && o.kind == PQC_OP_H) code = 2;
              if (cls == 2 && o.kind == PQC_OP_RZ && kof[o.b0] >= 0) code = 1;
              if (!code) continue;
              const int k = kof[o.b0];
              if (k < 0) return false;
              const bool is_merged = merged.count(i) != 0;
              // a merged ry carries its own frame update: it may not share a REAL4 op (which
              // flushes the frame for its Hadamards) -- the mix only occurs in initial layers
              if (is_merged && level_has_h) return false;
              size_t q = 0;
              while (q < packs.size() && ((packs[q].sub >> (2 * k)) & 3)) ++q;
              if (q == packs.size()) {
                TPOp t;
                memset(&t, 0, sizeof(t));
                t.kind = (uint8_t)(cls == 0 ? PQC_K_LAYER_RX4
                                            : (cls == 2 ? PQC_K_LAYER_RZ4
                                                        : (level_has_h ? PQC_K_LAYER_REAL4 : PQC_K_LAYER_RY4)));
                t.a = t.b = 0xff;
                if (t.kind == PQC_K_LAYER_RY4) t.wt = t.nterms = 0xff;
                packs.push_back(t);
              }
              packs[q].sub |= (uint8_t)(code << (2 * k));
              if (is_merged) {
                const std::pair<double, int>& mg = merged[i];
                packs[q].pad = 1;                  // the kernel's partner-aware path
                packs[q].t[k] = (uint16_t)trig_pair(o, mg.first);
                const uint8_t pb = partner_byte(mg.second);
                if (k == 0) packs[q].a = pb;
                else if (k == 1) packs[q].b = pb;
                else if (k == 2) packs[q].wt = pb;
                else packs[q].nterms = pb;
              } else if (code == 1) {
                packs[q].t[k] = (uint16_t)trig_of(0, o);
              }
            }
            for (const TPOp& t : packs)
              if (!push(t)) return false;
          }
          // CZ(register bit, thread-constant bit): the pending Z frame
          for (int i : L) {
            const FOp& o = F.ops[i];
            if (o.kind != PQC_OP_CZ) continue;
            const int k0 = kof[o.b0], k1 = kof[o.b1];
            if ((k0 >= 0) == (k1 >= 0) || no_czf) continue;
            TPOp t;
            memset(&t, 0, sizeof(t));
            t.kind = PQC_K_CZF;
            t.a = (uint8_t)(k0 >= 0 ? k0 : k1);
            t.b = partner_byte(k0 >= 0 ? o.b1 : o.b0);
            if (!push(t)) return false;
          }
          // XY pair rotations
          for (int i : L) {
            const FOp& o = F.ops[i];
            if (o.kind != PQC_K_RXY) continue;
            const int ka = std::min(kof[o.b0], kof[o.b1]), kb = std::max(kof[o.b0], kof[o.b1]);
            if (ka < 0) FRONT_FAIL(13);
            TPOp t;
            memset(&t, 0, sizeof(t));
            t.kind = PQC_K_RXY;
            t.a = (uint8_t)(ka * 4 + kb);
            t.b = 0xff;
            t.t[0] = (uint16_t)trig_of(2, o);
            if (!push(t)) FRONT_FAIL(14);
          }
          // thread-level R_z phases and CZ signs: one run, applied once
          for (int i : L) {
            const FOp& o = F.ops[i];
            if (o.kind == PQC_OP_RZ && kof[o.b0] < 0) {
              TPOp t;
              memset(&t, 0, sizeof(t));
              t.kind = PQC_OP_RZ;
              t.a = 0xff;
              t.b = lpos[o.b0] >= 0 ? (uint8_t)lpos[o.b0] : 0xff;
              t.t[0] = (uint16_t)trig_of(1, o);
              t.t[1] = (uint16_t)o.b0;
              if (!push(t)) FRONT_FAIL(15);
            } else if (o.kind == PQC_OP_CZ && (no_czf || (kof[o.b0] >= 0) == (kof[o.b1] >= 0))) {
              TPOp t;
              memset(&t, 0, sizeof(t));
              t.kind = PQC_OP_CZ;
              t.a = kof[o.b0] >= 0 ? (uint8_t)kof[o.b0] : 0xff;
              t.b = kof[o.b1] >= 0 ? (uint8_t)kof[o.b1] : 0xff;
              t.t[0] = lpos[o.b0] >= 0 ? (uint16_t)lpos[o.b0] : 0xffff;
              t.t[1] = lpos[o.b1] >= 0 ? (uint16_t)lpos[o.b1] : 0xffff;
              t.t[2] = (uint16_t)o.b0;
              t.t[3] = (uint16_t)o.b1;
              if (!push(t)) FRONT_FAIL(16);
            }
          }
          // same-angle R_zz of the level: one table-lookup phase per angle
          {
            std::vector<char> usedz(L.size(), 0);
            for (size_t x = 0; x < L.size(); ++x) {
              if (usedz[x] || F.ops[L[x]].kind != PQC_OP_RZZ) continue;
              std::vector<int> mem;
              for (size_t y = x; y < L.size(); ++y) {
                const FOp& o = F.ops[L[y]];
                if (usedz[y] || o.kind != PQC_OP_RZZ || !same_angle(o, F.ops[L[x]])) continue;
                bool dup = false;
                for (int m : mem) dup = dup || F.ops[m].support == o.support;
                if (dup || (int)mem.size() >= V1_MAX_TERMS) continue;
                mem.push_back(L[y]);
                usedz[y] = 1;
              }
              // one or two bonds that sit entirely on the register bits (the brick-wall diamonds
              // of the XXZ template, two thirds of its R_zz groups): a direct phase on the
              // thread's 16 amplitudes -- 64 / 32 FP64 instructions instead of a table-lookup op
              // (per amplitude LOP3 + POPC + LDS.128 + a complex product, ~165 in all)
              {
                static const bool no_rzzr = getenv("PQC_FRONT_NORZZR") != nullptr;
                bool regs = !no_rzzr && mem.size() <= 2;
                for (int m : mem) regs = regs && kof[F.ops[m].b0] >= 0 && kof[F.ops[m].b1] >= 0;
                if (regs && mem.size() == 2 && (F.ops[mem[0]].support & F.ops[mem[1]].support)) regs = false;
                if (regs) {
                  const FOp& o0 = F.ops[mem[0]];
                  const int ka = std::min(kof[o0.b0], kof[o0.b1]), kb = std::max(kof[o0.b0], kof[o0.b1]);
                  TPOp t;
                  memset(&t, 0, sizeof(t));
                  t.b = 0xff;
                  if (mem.size() == 1) {
                    t.kind = PQC_K_RZZ1;
                    t.a = (uint8_t)(ka * 4 + kb);
                    t.t[0] = (uint16_t)trig_of(1, o0);
                  } else {
                    // the pair that holds register bit 0 names the pairing: 01|23, 02|13, 03|12
                    const FOp& o1 = F.ops[mem[1]];
                    const int partner0 = ka == 0 ? kb : (kof[o1.b0] == 0 ? kof[o1.b1] : kof[o1.b0]);
                    t.kind = PQC_K_RZZ2;
                    t.a = (uint8_t)(partner0 - 1);
                    t.t[0] = (uint16_t)trig_of(2, o0);
                  }
                  if (!push(t)) FRONT_FAIL(18);
                  continue;
                }
              }
              if (nwt >= TP_MAX_WT) FRONT_FAIL(17);
              uint32_t wt[33];
              memset(wt, 0, sizeof(wt));
              for (size_t q = 0; q < mem.size(); ++q) {
                wt[F.ops[mem[q]].b0] ^= 1u << q;
                wt[F.ops[mem[q]].b1] ^= 1u << q;
              }
              for (int nib = 0; nib < 3; ++nib)
                for (int v = 0; v < 16; ++v) {
                  uint32_t w = 0;
                  for (int i2 = 0; i2 < 4; ++i2)
                    if ((v >> i2) & 1) w ^= wt[tb[4 * nib + i2]];
                  pp.wn[nwt][nib][v] = w;
                }
              for (int j = 0; j < n - 12; ++j) pp.wo[nwt][j] = wt[ps.obit[j]];
              TPOp t;
              memset(&t, 0, sizeof(t));
              t.kind = PQC_K_ZZSUM;
              t.a = t.b = 0xff;
              t.t[0] = (uint16_t)ntrig;
              t.wt = (uint8_t)nwt;
              t.nterms = (uint8_t)mem.size();
              const FOp& o0 = F.ops[mem[0]];
              for (int q = 0; q <= (int)mem.size(); ++q) {
                TrigJob j;
                memset(&j, 0, sizeof(j));
                j.kind = PQC_K_ZZSUM;
                j.param = o0.param;
                j.param2 = -1;
                j.slot = ntrig + q;
                j.npairs = (int)mem.size();
                j.pad = q;
                j.scale = o0.scale;
                j.offset = o0.offset;
                tj.push_back(j);
              }
              ntrig += (int)mem.size() + 1;
              ++nwt;
              if (!push(t)) FRONT_FAIL(18);
            }
          }
        }
        for (int i : sim.body)
          ps.op_ids.insert(ps.op_ids.end(), F.ops[i].ids.begin(), F.ops[i].ids.end());
      }
      for (int i : sim.post) {
        if (!push(perm_op(F.ops[i]))) FRONT_FAIL(19);
        ps.op_ids.insert(ps.op_ids.end(), F.ops[i].ids.begin(), F.ops[i].ids.end());
      }
      sw.npost = (uint8_t)sim.post.size();
      sw.op_end = (uint16_t)nops;
      if (sw.npost && !build_aff(nops - sw.npost, nops, false, sw.post, last)) FRONT_FAIL(22);
    }
    if (ntrig > TP_MAX_TRIG) FRONT_FAIL(20);
    pp.nops = nops;
    pp.ntrig = std::max(1, ntrig);
    pp.nwt = nwt;
    ps.nsweeps = pp.nsw;
    ps.nmops = nops;
    ps.sweep_off = ps.mop_off = 0;
    ps.tj_off = (int)tjobs.size();
    ps.ntjobs = (int)tj.size();
    ps.ntrig = pp.ntrig;
    ps.trig_goff = plan_slots;
    for (TrigJob j : tj) {
      j.slot += plan_slots;
      tjobs.push_back(j);
    }
    plan_slots += ps.ntrig;
    ps.wt_off = 0;
    ps.nwt = 0;
    ps.io_first = ps.io_last = 0;
    ps.pipe_idx = (int)prog->h_pipe.size();
    prog->h_pipe.push_back(pp);
    prog->v1_passes.push_back(ps);
    run_list.push_back((int)prog->v1_passes.size() - 1);
  }
  prog->front_run = run_list;
  prog->front_tj0 = (int)tj0;
  prog->front_ntj = (int)(tjobs.size() - tj0);
  prog->front_slots = plan_slots;
  prog->front_ok = true;
  return true;
}

bool pqc_plan_front(pqc_program* prog, std::vector<TrigJob>& tjobs) {
  const size_t pass0 = prog->v1_passes.size(), pipe0 = prog->h_pipe.size(), tj0 = tjobs.size();
  prog->front_ok = false;
  prog->front_run.clear();
  if (front_impl(prog, tjobs)) return true;
  prog->v1_passes.resize(pass0);
  prog->h_pipe.resize(pipe0);
  tjobs.resize(tj0);
  prog->front_run.clear();
  prog->front_ok = false;
  return false;
}
