// pqc_v1.cu -- the register-blocked "sweep" engine (v1) and its planner.
//
// Same job as pqc_apply.cu (circuit.py:123-124 gate application, circuit.py:149-192
// derivative states, measure.py:33-71 Gram entries) but organised for the B200 memory
// system:
//   * a PASS stages a 4096-amplitude tile (12 tile bits, the low 4 always included so
//     every global access is a full 256 B run) and visits it in SWEEPS;
//   * a SWEEP pulls 16 amplitudes per thread into registers (4 "register bits") and
//     applies every pending op on those bits plus every pending diagonal op -- one
//     shared-memory round trip and one barrier per sweep instead of one per gate;
//     the first / last sweep of a pass talk to global memory directly;
//   * shared memory is XOR-swizzled so the three aligned nibble sweeps are conflict free;
//   * the planner groups the op list into mutually commuting BLOCKS, fuses same-angle
//     R_zz sets into one table-lookup phase (ZZSUM) and R_yy*R_xx pairs into one XY
//     rotation, lets passes run across parameter boundaries, spawns derivative vectors
//     where the generator commutes with the rest of its block, multiplies diagonal
//     generators inside the pass (recomputing psi's tile), and takes the Gram dot
//     products while the next pass loads its tile (ping-pong buffers keep that race free).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "pqc_common.cuh"
#include "pqc_ops.cuh"


// Which plan runs PQC.run.  The front plan (pqc_front.cu, few deep passes on k_tile_pipe) wins
// wherever the block plan would run passes on the generic interpreter (CNOT-chain circuits:
// 2.8x), k_layer_seq passes with runs of R_z / CZ (NPQC: 1.6x - 4x), and -- since the register-bond
// R_zz are direct phases and k_tile_pipe has an instance of its own for the XXZ op family -- the XXZ
// template's layer-sequence passes at every size measured (12q x 16: 4.7 against 8.5 ms per 8192
// states, 16q x 16: 27.5 / 43.9 ms per 2048, 20q x 8: 37.1 / 55.2 ms per 256, 24q x 4: 26.3 / 33.5 ms
// per 16, 28q x 2: 35.8 / 41.0 ms per 2: profiles/r2_rzz_register_ops.md, r2_relabel_tables.md).
// The block plan stays for circuits without XY rotations whose passes are all layer passes or
// plain layer-sequence passes (TFIM: k_layer_pass' compile-time geometry, 1.7x).
// PQC_FRONT=0 / 1 forces the choice (read per call so tests can compare the plans).
bool pqc_use_front(const pqc_program* prog) {
  if (!prog->front_ok || prog->front_run.empty()) return false;
  const char* e = getenv("PQC_FRONT");
  if (e && strcmp(e, "0") == 0) return false;
  if (e && strcmp(e, "1") == 0) return true;
  // layer passes (compile-time geometry) and layer-sequence passes without R_z / CZ runs are the
  // light block-plan passes: they keep TFIM-type circuits (at 18+ qubits some of their passes need
  // another sweep order and run on k_layer_seq; the front plan is 1.7x slower on TFIM at 16 qubits).
  // A circuit with XY pair rotations (the XXZ type) takes the front plan although its passes are light.
  bool all_light = !prog->v1_run.empty();
  for (int pi : prog->v1_run) {
    const V1Pass& ps = prog->v1_passes[pi];
    all_light = all_light && (ps.fast_ok || (ps.seq_ok && !ps.seq.has_diag));
  }
  bool has_xy = false;
  for (const pqc_op& o : prog->ops) has_xy = has_xy || o.kind == PQC_OP_RXX || o.kind == PQC_OP_RYY;
  return !all_light || has_xy;
}

bool pqc_use_v0() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PQC_ENGINE");
    v = (e && strcmp(e, "v0") == 0) ? 1 : 0;
  }
  return v == 1;
}

// =====================================================================================
// planner
// =====================================================================================
namespace {

struct COp {
  int kind = 0, b0 = -1, b1 = -1, param = -1, param2 = -1, group = -1;
  double scale = 1.0, offset = 0.0;
  uint32_t support = 0, mix = 0;
  bool diag = false, pauli = false, two = false;
  uint32_t px = 0, pz = 0, px2 = 0, pz2 = 0;
  std::vector<std::pair<int, int>> pairs;   // ZZSUM
  std::vector<int> ids;                     // indices of the program ops this op stands for
};

bool pauli_comm(uint32_t x1, uint32_t z1, uint32_t x2, uint32_t z2) {
  return ((__builtin_popcount(x1 & z2) + __builtin_popcount(z1 & x2)) & 1) == 0;
}

bool commute(const COp& a, const COp& b) {
  if (!(a.support & b.support)) return true;
  if (a.diag && b.diag) return true;
  if (a.pauli && b.pauli) {
    bool ok = pauli_comm(a.px, a.pz, b.px, b.pz);
    if (a.two) ok = ok && pauli_comm(a.px2, a.pz2, b.px, b.pz);
    if (b.two) ok = ok && pauli_comm(a.px, a.pz, b.px2, b.pz2);
    if (a.two && b.two) ok = ok && pauli_comm(a.px2, a.pz2, b.px2, b.pz2);
    return ok;
  }
  return false;
}

COp make_cop(const pqc_op& op, int n, int idx) {
  COp c;
  c.ids.push_back(idx);
  c.kind = op.kind;
  c.b0 = n - 1 - op.q0;
  c.b1 = op.q1 >= 0 ? n - 1 - op.q1 : -1;
  c.param = op.param;
  c.param2 = op.param2;
  c.scale = op.scale;
  c.offset = op.offset;
  c.group = op.group;
  const uint32_t m0 = 1u << c.b0, m1 = c.b1 >= 0 ? 1u << c.b1 : 0u;
  c.support = m0 | m1;
  switch (op.kind) {
    case PQC_OP_RX: c.mix = m0; c.pauli = true; c.px = m0; break;
    case PQC_OP_RY: c.mix = m0; c.pauli = true; c.px = m0; c.pz = m0; break;
    case PQC_OP_RZ: c.diag = true; c.pauli = true; c.pz = m0; break;
    case PQC_OP_H: case PQC_OP_X: c.mix = m0; break;
    case PQC_OP_S: case PQC_OP_T: case PQC_OP_IDENT: c.diag = true; break;
    case PQC_OP_CNOT: c.mix = m1; break;
    case PQC_OP_CZ: c.diag = true; break;
    case PQC_OP_RXX: c.mix = m0 | m1; c.pauli = true; c.px = m0 | m1; break;
    case PQC_OP_RYY: c.mix = m0 | m1; c.pauli = true; c.px = m0 | m1; c.pz = m0 | m1; break;
    case PQC_OP_RZZ: c.diag = true; c.pauli = true; c.pz = m0 | m1; break;
    default: c.mix = m0 | m1; break;   // SQRTISWAP, FSIM, FIXED_FSIM
  }
  return c;
}

bool same_angle(const COp& a, const COp& b) {
  return a.param == b.param && a.scale == b.scale && a.offset == b.offset;
}

// fuse inside one reference gate (same `group`) when all its members commute
void fuse_group(std::vector<COp>& run) {
  if (run.size() < 2) return;
  for (auto& a : run)
    if (!a.pauli) return;
  for (size_t i = 0; i < run.size(); ++i)
    for (size_t j = i + 1; j < run.size(); ++j)
      if (!commute(run[i], run[j])) return;
  std::vector<bool> dead(run.size(), false);
  // same-angle R_zz sets -> ZZSUM
  for (size_t i = 0; i < run.size(); ++i) {
    if (dead[i] || run[i].kind != PQC_OP_RZZ) continue;
    std::vector<size_t> mem;
    for (size_t j = i; j < run.size(); ++j)
      if (!dead[j] && run[j].kind == PQC_OP_RZZ && same_angle(run[i], run[j])) mem.push_back(j);
    if (mem.size() < 2 || mem.size() > V1_MAX_TERMS) continue;
    bool dup = false;                      // RR_block on 2 qubits lists its pair twice (quirk Q9)
    for (size_t u = 0; u < mem.size(); ++u)
      for (size_t v = u + 1; v < mem.size(); ++v)
        if (run[mem[u]].support == run[mem[v]].support) dup = true;
    if (dup) continue;
    COp z = run[i];
    z.kind = PQC_K_ZZSUM;
    z.pauli = false;
    z.diag = true;
    z.mix = 0;
    z.support = 0;
    z.ids.clear();
    for (size_t j : mem) {
      z.pairs.push_back({run[j].b0, run[j].b1});
      z.support |= run[j].support;
      z.ids.insert(z.ids.end(), run[j].ids.begin(), run[j].ids.end());
      if (j != i) dead[j] = true;
    }
    run[i] = z;
  }
  // R_yy * R_xx on the same pair with the same angle -> RXY
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
  std::vector<COp> out;
  for (size_t i = 0; i < run.size(); ++i)
    if (!dead[i]) out.push_back(run[i]);
  run.swap(out);
}

int trig_entries(const COp& c) {
  switch (c.kind) {
    case PQC_OP_RX: case PQC_OP_RY: case PQC_OP_RZ: case PQC_OP_RXX: case PQC_OP_RYY:
    case PQC_OP_RZZ: case PQC_K_RXY: return 1;
    case PQC_OP_FSIM: case PQC_OP_FIXED_FSIM: return 2;
    case PQC_K_ZZSUM: return (int)c.pairs.size() + 1;
    default: return 0;
  }
}

struct Builder {
  const pqc_program* prog;
  int n, cap, ipc;                         // qubits, tile amplitude bits, items per CTA
  std::vector<MOp>& mops;
  std::vector<SweepD>& sweeps;
  std::vector<TrigJob>& tjobs;
  std::vector<uint32_t>& zz;
  std::vector<V1Pass>& passes;

  // current pass
  uint32_t tile = 0;
  std::vector<int> grp[3];
  int grp_of[32];
  // pre/post: index permutations (X, CNOT) folded into the sweep's load / store addresses
  struct SW { int g; std::vector<MOp> pre, ops, post; };
  std::vector<SW> sws;
  std::vector<TrigJob> tj;
  int ntrig = 0;
  std::vector<int> spawn_param;
  bool open = false;
  int nmops_cur = 0;
  std::vector<int> cur_ids;                // program ops taken by the current pass
  int wt_begin = 0;                        // first linear-form table of the current pass
  int plan_slots = 0;                      // trig slots used by the finished passes of this plan

  uint32_t lowmask() const { return n >= 4 ? 0xfu : ((1u << n) - 1u); }

  void begin() {
    tile = 0;
    for (auto& g : grp) g.clear();
    for (int& g : grp_of) g = -1;
    sws.clear();
    cur_ids.clear();
    tj.clear();
    ntrig = 0;
    spawn_param.clear();
    open = true;
    nmops_cur = 0;
    wt_begin = (int)zz.size();
  }
  int n_tables() const { return ((int)zz.size() - wt_begin) / V1_WTAB; }
  int room(int g, uint32_t need) const {   // free slots of group g for bits of `need`
    int r = 4 - (int)grp[g].size();
    if (g == 0) {                          // keep space for low bits that are still unplaced
      int res = 0;
      for (int b = 0; b < 4 && b < n; ++b) res += (grp_of[b] < 0) && !((need >> b) & 1u);
      r -= res;
    }
    return r;
  }
  // group an op with mixing bits `need` would go to; -1 if it does not fit this pass
  int target_group(uint32_t need) const {
    if (__builtin_popcount(tile | need | lowmask()) > cap) return -1;
    int g = -1, fresh = 0;
    for (int b = 0; b < n; ++b)
      if ((need >> b) & 1u) {
        if (grp_of[b] >= 0) {
          if (g >= 0 && g != grp_of[b]) return -1;
          g = grp_of[b];
        } else {
          ++fresh;
        }
      }
    if (g >= 0) return room(g, need) >= fresh ? g : -1;
    const int cur = sws.empty() ? -1 : sws.back().g;
    if (cur >= 0 && room(cur, need) >= fresh) return cur;
    const bool all_low = (need & ~lowmask()) == 0;
    const int order_low[3] = {0, 1, 2}, order_hi[3] = {2, 1, 0};
    for (int i = 0; i < 3; ++i) {
      const int c = all_low ? order_low[i] : order_hi[i];
      if (room(c, need) >= fresh) return c;
    }
    return -1;
  }
  bool trig_fits(const COp& c) const { return (ntrig + trig_entries(c)) * ipc <= 2048; }

  int priority(const COp& c) const {       // lower is better; 99 = cannot take now
    if (!trig_fits(c) || sws.size() >= 30 || nmops_cur >= 150) return 99;
    if (c.kind == PQC_K_ZZSUM && n_tables() >= V1_MAX_WT) return 99;
    if (!c.mix) return 0;
    const int g = target_group(c.mix);
    if (g < 0) return 99;
    const bool perm = c.kind == PQC_OP_X || c.kind == PQC_OP_CNOT;
    if (!sws.empty() && (sws.back().g == g || sws.back().g < 0) &&
        (perm || sws.back().post.empty()))
      return 1;
    bool assigned = false;
    for (int b = 0; b < n; ++b)
      if (((c.mix >> b) & 1u) && grp_of[b] >= 0) assigned = true;
    if (assigned) return 2;
    // fresh bits: open the pass on high bits (direct global load), take the low nibble in the
    // middle, finish on high bits again (direct global store)
    const bool all_low = (c.mix & ~lowmask()) == 0;
    const bool want_low = !sws.empty() && grp[0].empty();
    return (all_low == want_low) ? 3 : 4;
  }

  MOp proto(const COp& c) {
    MOp m;
    memset(&m, 0, sizeof(m));
    m.kind = c.kind;
    m.b0 = c.b0;
    m.b1 = c.b1;
    m.k0 = m.k1 = m.l0 = m.l1 = -1;
    m.trig = -1;
    const int te = trig_entries(c);
    if (te && c.kind != PQC_K_ZZSUM) {
      // ops driven by the same angle (shared_parameter layers) share one trig slot
      const int want_pad = (c.kind == PQC_OP_RX || c.kind == PQC_OP_RY) ? 1 : 0;
      for (const TrigJob& e : tj) {
        const bool same_class = (e.kind == c.kind) ||
                                (e.pad == 1 && want_pad == 1) ||
                                (e.pad == 0 && want_pad == 0 && e.kind != PQC_K_ZZSUM &&
                                 e.kind != PQC_OP_FSIM && e.kind != PQC_OP_FIXED_FSIM &&
                                 e.kind != PQC_K_RXY && c.kind != PQC_OP_FSIM &&
                                 c.kind != PQC_OP_FIXED_FSIM && c.kind != PQC_K_RXY);
        if (same_class && e.param == c.param && e.param2 == c.param2 && e.scale == c.scale &&
            e.offset == c.offset && e.npairs == 0) {
          m.trig = e.slot;
          return m;
        }
      }
    }
    if (te) {
      m.trig = ntrig;
      TrigJob j;
      memset(&j, 0, sizeof(j));
      j.kind = c.kind;
      j.param = c.param;
      j.param2 = c.param2;
      j.slot = ntrig;
      j.npairs = (int)c.pairs.size();
      j.scale = c.scale;
      j.offset = c.offset;
      if (c.kind == PQC_OP_RX || c.kind == PQC_OP_RY) j.pad = 1;   // stored as (tan, cos)
      if (c.kind == PQC_K_ZZSUM) {          // one job per table entry so they fill in parallel
        for (int k = 0; k < te; ++k) {
          j.slot = ntrig + k;
          j.pad = k;
          tj.push_back(j);
        }
      } else {
        tj.push_back(j);
      }
      ntrig += te;
    }
    if (c.kind == PQC_K_ZZSUM) {
      // linear-form table: bit k of word b <=> index bit b belongs to pair k
      m.npairs = (int)c.pairs.size();
      m.aux1 = (int)zz.size() - wt_begin;
      zz.resize(zz.size() + V1_WTAB, 0u);
      uint32_t* w = zz.data() + wt_begin + m.aux1;
      for (size_t k = 0; k < c.pairs.size(); ++k) {
        w[c.pairs[k].first] ^= 1u << k;
        w[c.pairs[k].second] ^= 1u << k;
      }
    }
    return m;
  }

  void take(const COp& c) {
    MOp m = proto(c);
    cur_ids.insert(cur_ids.end(), c.ids.begin(), c.ids.end());
    ++nmops_cur;
    if (!c.mix) {
      if (sws.empty() || !sws.back().post.empty()) sws.push_back(SW{-1, {}, {}, {}});
      sws.back().ops.push_back(m);
      return;
    }
    const bool perm = c.kind == PQC_OP_X || c.kind == PQC_OP_CNOT;
    const int g = target_group(c.mix);
    for (int b = 0; b < n; ++b)
      if (((c.mix >> b) & 1u) && grp_of[b] < 0) {
        grp_of[b] = g;
        grp[g].push_back(b);
      }
    tile |= c.mix;
    if (!sws.empty() && sws.back().g < 0) sws.back().g = g;
    if (sws.empty() || sws.back().g != g || (!perm && !sws.back().post.empty()))
      sws.push_back(SW{g, {}, {}, {}});
    SW& sw = sws.back();
    if (!perm) sw.ops.push_back(m);
    else if (sw.ops.empty() && sw.post.empty()) sw.pre.push_back(m);
    else sw.post.push_back(m);
  }

  void take_gen(int param) {               // in-pass diagonal generator multiply
    MOp m;
    memset(&m, 0, sizeof(m));
    m.kind = PQC_K_GEN;
    m.k0 = m.k1 = m.l0 = m.l1 = m.b0 = m.b1 = -1;
    m.trig = -1;
    m.aux0 = (int)spawn_param.size();
    // linear-form table of the generator's z-masks: bit t of word b = bit b of z_t
    m.aux1 = (int)zz.size() - wt_begin;
    zz.resize(zz.size() + V1_WTAB, 0u);
    uint32_t* w = zz.data() + wt_begin + m.aux1;
    const int g0 = prog->gen_off[param], g1 = prog->gen_off[param + 1];
    m.npairs = g1 - g0;
    for (int t = g0; t < g1; ++t)
      for (int b = 0; b < 32; ++b)
        if ((prog->gens[t].zmask >> b) & 1u) w[b] |= 1u << (t - g0);
    spawn_param.push_back(param);
    ++nmops_cur;
    if (sws.empty() || !sws.back().post.empty()) sws.push_back(SW{-1, {}, {}, {}});
    sws.back().ops.push_back(m);
  }

  int close() {                            // -> index of the finished pass
    // place the low bits, then fill the tile with the lowest unused global bits
    auto place = [&](int b) {
      const int pref[3] = {0, 1, 2};
      for (int i = 0; i < 3; ++i)
        if ((int)grp[pref[i]].size() < 4) {
          grp_of[b] = pref[i];
          grp[pref[i]].push_back(b);
          tile |= 1u << b;
          return;
        }
    };
    for (int b = 0; b < 4 && b < n; ++b)
      if (grp_of[b] < 0) place(b);
    for (int b = 0; b < n && __builtin_popcount(tile) < cap; ++b)
      if (grp_of[b] < 0) place(b);
    V1Pass ps;
    memset(ps.lbit, 0, sizeof(ps.lbit));
    memset(ps.obit, 0, sizeof(ps.obit));
    ps.tb = cap;
    int lpos[32];
    for (int& v : lpos) v = -1;
    int pos = 0;
    int gfirst[3], gcount[3];
    for (int g = 0; g < 3; ++g) {
      std::sort(grp[g].begin(), grp[g].end());
      gfirst[g] = pos;
      gcount[g] = (int)grp[g].size();
      for (int b : grp[g]) {
        lpos[b] = pos;
        ps.lbit[pos++] = b;
      }
    }
    int o = 0;
    for (int b = 0; b < n; ++b)
      if (lpos[b] < 0) ps.obit[o++] = b;
    ps.low_run = 0;
    while (ps.low_run < cap && ps.lbit[ps.low_run] == ps.low_run) ps.low_run++;
    // a sweep may talk to global memory directly when the three lowest index bits are lane
    // bits: every 8 lanes then cover one full 128-byte line
    ps.direct_ok = (n >= V1_LOCAL_BITS) && ps.low_run >= 3;
    if (sws.empty()) sws.push_back(SW{-1, {}, {}, {}});
    ps.sweep_off = (int)sweeps.size();
    ps.mop_off = (int)mops.size();
    for (size_t si = 0; si < sws.size(); ++si) {
      SW& sw = sws[si];
      if (sw.g < 0) sw.g = gcount[2] ? 2 : (gcount[1] ? 1 : 0);
      SweepD d;
      memset(&d, 0, sizeof(d));
      std::vector<int> rb;
      for (int b : grp[sw.g]) rb.push_back(lpos[b]);
      // fewer than 4 bits in the group (n < 12): borrow neighbouring amplitude positions
      for (int g2 = 1; g2 <= 2 && rb.size() < 4; ++g2) {
        const int gg = (sw.g + g2) % 3;
        for (int i = gcount[gg] - 1; i >= 0 && rb.size() < 4; --i) rb.push_back(gfirst[gg] + i);
      }
      std::sort(rb.begin(), rb.end());
      for (int k = 0; k < 4; ++k) d.rb[k] = rb[k];
      for (int k = 0; k < 4; ++k) {
        const unsigned i = 1u << rb[k];
        d.sz[k] = (unsigned short)(i ^ (((i >> 3) ^ (i >> 6) ^ (i >> 9)) & 7u));
      }
      {
        int t = 0;
        for (int posn = 0; posn < V1_LOCAL_BITS; ++posn) {
          if (std::find(rb.begin(), rb.end(), posn) != rb.end()) continue;
          d.tl[t] = (unsigned char)posn;
          d.tg[t] = posn < cap ? (unsigned char)ps.lbit[posn] : (unsigned char)255;
          ++t;
        }
        for (int k = 0; k < 4; ++k) {
          d.gb[k] = (unsigned char)ps.lbit[rb[k]];
          d.gm[k] = 1u << ps.lbit[rb[k]];
        }
        for (int h = 0; h < 2; ++h)
          for (int v = 0; v < 16; ++v) {
            unsigned pos = 0, amp = 0;
            for (int i = 0; i < 4 && 4 * h + i < V1_TBITS; ++i)
              if ((v >> i) & 1) {
                pos |= 1u << d.tl[4 * h + i];
                if (d.tg[4 * h + i] < 32) amp |= 1u << d.tg[4 * h + i];
              }
            d.tpos[h][v] = (unsigned short)pos;
            d.tamp[h][v] = amp;
          }
      }
      d.mop_begin = (int)mops.size() - ps.mop_off;
      auto emit = [&](std::vector<MOp>& list) {
        for (MOp m : list) {
          auto fill = [&](int b, int& k, int& l) {
            if (b < 0) return;
            l = lpos[b];
            k = -1;
            for (int q = 0; q < 4; ++q)
              if (l >= 0 && d.rb[q] == l) k = q;
          };
          fill(m.b0, m.k0, m.l0);
          fill(m.b1, m.k1, m.l1);
          mops.push_back(m);
        }
      };
      emit(sw.pre);
      {
        // resolve register-bit indices, then pack runs of RX / RY / H into LAYER1Q macro-ops
        const size_t at = mops.size();
        emit(sw.ops);
        std::vector<MOp> packed;
        for (size_t q = at; q < mops.size(); ++q) {
          const MOp& m = mops[q];
          const bool oneq = m.kind == PQC_OP_RX || m.kind == PQC_OP_RY || m.kind == PQC_OP_H;
          if (!oneq) { packed.push_back(m); continue; }
          const int lk = m.kind == PQC_OP_RX ? PQC_K_LAYER_RX4 : PQC_K_LAYER_REAL4;
          if (packed.empty() || packed.back().kind != lk ||
              ((packed.back().subk >> (8 * m.k0)) & 0xff)) {
            MOp L;
            memset(&L, 0, sizeof(L));
            L.kind = lk;
            L.k0 = L.k1 = L.l0 = L.l1 = L.b0 = L.b1 = -1;
            L.trig = -1;
            packed.push_back(L);
          }
          packed.back().subk |= (m.kind + 1) << (8 * m.k0);
          packed.back().subt[m.k0] = m.trig < 0 ? 0 : m.trig;
        }
        mops.resize(at);
        for (auto& m : packed) mops.push_back(m);
      }
      emit(sw.post);
      d.mop_end = (int)mops.size() - ps.mop_off;
      d.pad = (int)sw.pre.size() | ((int)sw.post.size() << 16);
      const bool touches_low = d.rb[0] < 3;
      if (ps.direct_ok && !touches_low) {
        if (si == 0) d.io |= 1;
        if (si + 1 == sws.size()) d.io |= 2;
      }
      sweeps.push_back(d);
    }
    ps.nsweeps = (int)sws.size();
    ps.nmops = (int)mops.size() - ps.mop_off;
    ps.io_first = sweeps[ps.sweep_off].io;
    ps.io_last = sweeps[ps.sweep_off + ps.nsweeps - 1].io;
    ps.tj_off = (int)tjobs.size();
    ps.ntjobs = (int)tj.size();
    ps.ntrig = std::max(1, ntrig);
    ps.trig_goff = plan_slots;
    for (auto j : tj) {
      j.slot += plan_slots;                 // slot in the plan's per-sample table
      tjobs.push_back(j);
    }
    plan_slots += ps.ntrig;
    ps.wt_off = wt_begin;
    ps.nwt = n_tables();
    // ---- layer-pass fast path eligibility -------------------------------------------------
    ps.fast_ok = false;
    memset(&ps.fast, 0, sizeof(ps.fast));
    if (cap == V1_LOCAL_BITS && V1_LOCAL_BITS == 12 && ps.direct_ok && ps.low_run >= 4 &&
        ps.nsweeps >= 1 && ps.nsweeps <= 3 && (int)spawn_param.size() <= FAST_MAX_SPAWN &&
        (ps.io_first & 1) && (ps.io_last & 2)) {
      const int want3[3][4] = {{8, 9, 10, 11}, {0, 1, 2, 3}, {4, 5, 6, 7}};
      const int want2[2][4] = {{8, 9, 10, 11}, {4, 5, 6, 7}};
      bool ok = true;
      ps.fast.ns = ps.nsweeps;
      for (int si = 0; si < ps.nsweeps && ok; ++si) {
        const SweepD& d = sweeps[ps.sweep_off + si];
        const int* w = ps.nsweeps == 3 ? want3[si] : want2[si];
        for (int k = 0; k < 4; ++k) ok = ok && d.rb[k] == w[k];
        ok = ok && d.pad == 0 && d.mop_end - d.mop_begin <= FAST_MAX_OPS;
        const int slot = ps.nsweeps == 3 ? si : (si == 0 ? 0 : 2);
        ps.fast.nops[slot] = 0;
        for (int mi = d.mop_begin; mi < d.mop_end && ok; ++mi) {
          const MOp& m = mops[ps.mop_off + mi];
          FastOp f;
          memset(&f, 0, sizeof(f));
          f.kind = m.kind;
          f.subk = m.subk;
          if (m.kind == PQC_K_LAYER_RX4 || m.kind == PQC_K_LAYER_REAL4) {
            for (int k = 0; k < 4; ++k) f.t[k] = m.subt[k];
          } else if (m.kind == PQC_K_ZZSUM) {
            f.t[0] = m.trig;
            f.wt = m.aux1 / V1_WTAB;
            f.nterms = m.npairs;
          } else if (m.kind == PQC_K_GEN) {
            f.wt = m.aux1 / V1_WTAB;
            f.nterms = m.npairs;
            f.spawn = m.aux0;
          } else {
            ok = false;
          }
          ps.fast.ops[slot][ps.fast.nops[slot]++] = f;
        }
      }
      ok = ok && ps.nwt <= FAST_MAX_WT;
      if (ok) {
        for (int tb = 0; tb < ps.nwt; ++tb) {
          const uint32_t* wt = zz.data() + ps.wt_off + tb * V1_WTAB;
          for (int nib = 0; nib < 3; ++nib)
            for (int v = 0; v < 16; ++v) {
              uint32_t w = 0;
              for (int i = 0; i < 4; ++i)
                if ((v >> i) & 1) w ^= wt[ps.lbit[4 * nib + i]];
              ps.fast.wn[tb][nib][v] = w;
            }
          for (int j = 0; j < n - V1_LOCAL_BITS; ++j) ps.fast.wo[tb][j] = wt[ps.obit[j]];
        }
      }
      ps.fast_ok = ok;
    }
    // ---- layer-sequence path eligibility (any order of aligned nibble sweeps, XY pair ops) --
    ps.seq_ok = false;
    memset(&ps.seq, 0, sizeof(ps.seq));
    if (!ps.fast_ok && cap == V1_LOCAL_BITS && V1_LOCAL_BITS == 12 && n >= V1_LOCAL_BITS &&
        ps.nsweeps >= 1 && ps.nsweeps <= SEQ_MAX_SWEEPS &&
        (int)spawn_param.size() <= SEQ_MAX_SPAWN && ps.nwt <= FAST_MAX_WT) {
      bool ok = true;
      int nflat = 0;
      ps.seq.nsw = ps.nsweeps;
      for (int si = 0; si < ps.nsweeps && ok; ++si) {
        const SweepD& d = sweeps[ps.sweep_off + si];
        int g = -1;
        if (d.rb[0] == 8) g = 0; else if (d.rb[0] == 0) g = 1; else if (d.rb[0] == 4) g = 2;
        ok = ok && g >= 0;
        for (int k = 0; k < 4 && ok; ++k) ok = d.rb[k] == d.rb[0] + k;
        ok = ok && d.pad == 0 && nflat + d.mop_end - d.mop_begin <= SEQ_MAX_OPS;
        if (!ok) break;
        ps.seq.geom[si] = g;
        ps.seq.off[si] = nflat;
        ps.seq.nops[si] = 0;
        for (int mi = d.mop_begin; mi < d.mop_end && ok; ++mi) {
          const MOp& m = mops[ps.mop_off + mi];
          FastOp f;
          memset(&f, 0, sizeof(f));
          f.kind = m.kind;
          f.subk = m.subk;
          if (m.kind == PQC_K_LAYER_RX4 || m.kind == PQC_K_LAYER_REAL4) {
            for (int k = 0; k < 4; ++k) f.t[k] = m.subt[k];
          } else if (m.kind == PQC_K_ZZSUM) {
            f.t[0] = m.trig;
            f.wt = m.aux1 / V1_WTAB;
            f.nterms = m.npairs;
          } else if (m.kind == PQC_K_GEN) {
            f.wt = m.aux1 / V1_WTAB;
            f.nterms = m.npairs;
            f.spawn = m.aux0;
          } else if (m.kind == PQC_K_RXY && m.k0 >= 0 && m.k1 >= 0 && m.k0 != m.k1) {
            f.t[0] = m.trig;
            f.subk = std::min(m.k0, m.k1) * 4 + std::max(m.k0, m.k1);
          } else if (m.kind == PQC_OP_RZ) {
            ps.seq.has_diag = 1;
            f.t[0] = m.trig;
            f.t[1] = m.k0;
            f.t[2] = m.l0;
            f.t[3] = m.b0;
          } else if (m.kind == PQC_OP_CZ) {
            ps.seq.has_diag = 1;
            f.t[0] = m.k0;
            f.t[1] = m.k1;
            f.t[2] = m.l0;
            f.t[3] = m.l1;
            f.wt = m.b0;
            f.nterms = m.b1;
          } else if (m.kind == PQC_OP_IDENT) {
            continue;
          } else {
            ok = false;
          }
          ps.seq.ops[nflat++] = f;
          ps.seq.nops[si]++;
        }
      }
      if (ok) {
        for (int tb = 0; tb < ps.nwt; ++tb) {
          const uint32_t* wt = zz.data() + ps.wt_off + tb * V1_WTAB;
          for (int nib = 0; nib < 3; ++nib)
            for (int v = 0; v < 16; ++v) {
              uint32_t w = 0;
              for (int i = 0; i < 4; ++i)
                if ((v >> i) & 1) w ^= wt[ps.lbit[4 * nib + i]];
              ps.seq.wn[tb][nib][v] = w;
            }
          for (int j = 0; j < n - V1_LOCAL_BITS; ++j) ps.seq.wo[tb][j] = wt[ps.obit[j]];
        }
      }
      ps.seq_ok = ok;
    }
    ps.spawn_param = spawn_param;
    ps.op_ids = cur_ids;
    passes.push_back(ps);
    open = false;
    return (int)passes.size() - 1;
  }
};

}  // namespace

int pqc_plan_v1(pqc_program* prog) {
  const int n = prog->n, P = prog->P;
  prog->v1_ok = prog->v1_grad_ok = false;
  if (n < 8) return 0;                      // tiny circuits stay on the v0 tile kernel
  // ---- canonical ops: per-gate fusion --------------------------------------------------
  std::vector<COp> cops;
  {
    std::vector<COp> run;
    int cur = -2;
    auto flush = [&]() {
      fuse_group(run);
      for (auto& c : run) cops.push_back(c);
      run.clear();
    };
    for (size_t oi = 0; oi < prog->ops.size(); ++oi) {
      const pqc_op& op = prog->ops[oi];
      if (op.group != cur) {
        flush();
        cur = op.group;
      }
      run.push_back(make_cop(op, n, (int)oi));
    }
    flush();
  }
  // ---- mutually commuting blocks ----------------------------------------------------------
  std::vector<int> block_of(cops.size(), 0);
  std::vector<std::vector<int>> blocks;
  for (size_t i = 0; i < cops.size(); ++i) {
    bool ok = !blocks.empty();
    if (ok)
      for (int j : blocks.back())
        if (!commute(cops[i], cops[j])) { ok = false; break; }
    if (!ok) blocks.push_back({});
    blocks.back().push_back((int)i);
    block_of[i] = (int)blocks.size() - 1;
  }
  // ---- parameters: block and generator type -------------------------------------------------
  std::vector<int> param_block(P, -1);
  bool grad_ok = prog->grad_supported;
  for (size_t i = 0; i < cops.size() && grad_ok; ++i)
    for (int p : {cops[i].param, cops[i].param2})
      if (p >= 0) {
        if (param_block[p] < 0) param_block[p] = block_of[i];
        else if (param_block[p] != block_of[i]) grad_ok = false;   // spans blocks: keep v0
      }
  for (int p = 0; p < P && grad_ok; ++p)
    if (param_block[p] < 0) grad_ok = false;
  for (int p = 1; p < P && grad_ok; ++p)
    if (param_block[p] < param_block[p - 1]) grad_ok = false;
  std::vector<bool> param_diag(P, true);
  if (grad_ok)
    for (int p = 0; p < P; ++p) {
      if (prog->pspawn[p].type == 1) param_diag[p] = false;      // fSim family: pair-matrix spawn
      double cr = 0.0, ci = 0.0;
      for (int t = prog->gen_off[p]; t < prog->gen_off[p + 1]; ++t) {
        if (prog->gens[t].xmask) param_diag[p] = false;
        // the in-pass generator multiply assumes one common coefficient
        if (t > prog->gen_off[p] && (prog->gens[t].re != cr || prog->gens[t].im != ci))
          param_diag[p] = false;
        cr = prog->gens[t].re;
        ci = prog->gens[t].im;
      }
      if (prog->gen_off[p + 1] == prog->gen_off[p]) param_diag[p] = false;
      if (prog->gen_off[p + 1] - prog->gen_off[p] > V1_MAX_TERMS) param_diag[p] = false;
    }

  std::vector<MOp> mops;
  std::vector<SweepD> sweeps;
  std::vector<TrigJob> tjobs;
  std::vector<uint32_t> zz;
  const int cap = std::min(n, V1_LOCAL_BITS);
  const int ipc = 1 << (V1_LOCAL_BITS - cap);

  // one planning routine, with or without derivative markers
  auto plan = [&](bool markers, std::vector<int>* run_out, std::vector<V1Stage>* stages) -> int {
    Builder B{prog, n, cap, ipc, mops, sweeps, tjobs, zz, prog->v1_passes};
    std::vector<bool> spawned(P, false);
    for (int p = std::max(0, P - prog->n_nodiff); p < P; ++p) spawned[p] = true;   // angle only
    std::vector<int> pending_dots;
    const bool onload = (ipc == 1);
    auto emit_pass = [&]() {
      const int idx = B.close();
      if (run_out) run_out->push_back(idx);
      if (stages) {
        V1Stage st;
        st.type = 0;
        st.pass = idx;
        if (onload) {
          while (!pending_dots.empty() && (int)st.partners.size() < V1_MAX_PART) {
            st.partners.push_back(pending_dots.front());
            pending_dots.erase(pending_dots.begin());
          }
        }
        if (!pending_dots.empty()) {            // overflow or packed items: standalone dots first
          V1Stage d;
          d.type = 2;
          d.partners = pending_dots;
          pending_dots.clear();
          stages->push_back(d);
        }
        stages->push_back(st);
        for (int p : prog->v1_passes[idx].spawn_param) pending_dots.push_back(p);
      }
    };
    auto gather = [&](const std::vector<int>& ps) {
      if (ps.empty() || !stages) return;
      V1Stage st;
      st.type = 1;
      st.gather_params = ps;
      stages->push_back(st);
      for (int p : ps) pending_dots.push_back(p);
    };
    for (size_t bi = 0; bi < blocks.size(); ++bi) {
      std::vector<int> todo = blocks[bi];
      while (!todo.empty()) {
        if (!B.open) B.begin();
        int best = -1, bestp = 99;
        for (size_t k = 0; k < todo.size(); ++k) {
          const int pr = B.priority(cops[todo[k]]);
          if (pr < bestp) { bestp = pr; best = (int)k; }
          if (pr == 0) break;
        }
        if (best < 0) {
          // natural boundary inside the block: every not-yet-spawned parameter of this block
          // may be spawned here because its generator commutes with the rest of the block
          emit_pass();
          if (markers) {
            std::vector<int> ps;
            for (int p = 0; p < P; ++p)
              if (param_block[p] == (int)bi && !spawned[p]) { ps.push_back(p); spawned[p] = true; }
            gather(ps);
          }
          continue;
        }
        B.take(cops[todo[best]]);
        todo.erase(todo.begin() + best);
      }
      if (markers) {
        // block exhausted: diagonal generators are multiplied inside the pass ...
        std::vector<int> nd;
        for (int p = 0; p < P; ++p)
          if (param_block[p] == (int)bi && !spawned[p]) {
            if (param_diag[p] && onload) {
              if (!B.open) B.begin();
              if ((int)B.spawn_param.size() >= V1_MAX_SPAWN || B.n_tables() >= V1_MAX_WT) {
                emit_pass();
                B.begin();
              }
              B.take_gen(p);
              spawned[p] = true;
            } else {
              nd.push_back(p);
            }
          }
        // ... the others need the complete state: close the pass and gather
        if (!nd.empty()) {
          if (B.open) emit_pass();
          for (int p : nd) spawned[p] = true;
          gather(nd);
        }
      }
    }
    if (B.open || (run_out && run_out->empty())) {
      if (!B.open) B.begin();
      emit_pass();
    }
    if (stages && !pending_dots.empty()) {
      V1Stage d;
      d.type = 2;
      d.partners = pending_dots;
      stages->push_back(d);
    }
    return B.plan_slots;
  };

  prog->v1_run_tj0 = 0;
  prog->v1_run_slots = plan(false, &prog->v1_run, nullptr);
  prog->v1_run_ntj = (int)tjobs.size();
  prog->v1_ok = true;
  if (grad_ok && P > 0) {
    prog->v1_grad_tj0 = (int)tjobs.size();
    prog->v1_grad_slots = plan(true, nullptr, &prog->v1_grad);
    prog->v1_grad_ntj = (int)tjobs.size() - prog->v1_grad_tj0;
    prog->v1_grad_ok = true;
  }

  // tile-pipe form of every pass that has one (pqc_pipe.cu)
  prog->h_pipe.clear();
  for (V1Pass& ps : prog->v1_passes) {
    PipePlan pp;
    ps.pipe_idx = -1;
    if (pqc_pipe_build(ps, n, pp)) {
      ps.pipe_idx = (int)prog->h_pipe.size();
      prog->h_pipe.push_back(pp);
    }
  }
  // light-cone plan of PQC.run for k_tile_pipe (pqc_front.cu); appends passes and trig jobs
  pqc_plan_front(prog, tjobs);
  prog->h_mops = mops;
  prog->h_sweeps = sweeps;
  prog->h_tjobs = tjobs;
  prog->h_zz = zz;
  return 0;
}

// Human-readable plan (planner tests): one line per stage of the derivative plan
// (or of the run plan when there is none).
extern "C" PQC_API int pqc_program_describe(const pqc_program* prog, char* out, int64_t cap) {
  if (!prog || !out || cap <= 0) return -1;
  std::string s;
  char buf[256];
  auto pass_line = [&](const V1Pass& ps) {
    char buf[256];
    int nm = 0;
    std::string kinds;
    if (ps.front) {
      const PipePlan& pp = prog->h_pipe[ps.pipe_idx];
      for (int i = 0; i < pp.nsw; ++i) {
        const TPSweep& sw = pp.sw[i];
        snprintf(buf, sizeof(buf), " [rb %d,%d,%d,%d pre%d post%d:", sw.rpos[0], sw.rpos[1], sw.rpos[2],
                 sw.rpos[3], sw.npre, sw.npost);
        kinds += buf;
        for (int m = sw.op_begin; m < sw.op_end; ++m) {
          snprintf(buf, sizeof(buf), " %d", pp.ops[m].kind);
          kinds += buf;
        }
        kinds += "]";
      }
      snprintf(buf, sizeof(buf), "sweeps=%d ops=%d trig=%d tables=%d front=1", pp.nsw, pp.nops, pp.ntrig,
               pp.nwt);
      return std::string(buf) + kinds;
    }
    for (int i = 0; i < ps.nsweeps; ++i) {
      const SweepD& sw = prog->h_sweeps[ps.sweep_off + i];
      snprintf(buf, sizeof(buf), " [rb %d,%d,%d,%d io%d:", sw.rb[0], sw.rb[1], sw.rb[2], sw.rb[3], sw.io);
      kinds += buf;
      for (int m = sw.mop_begin; m < sw.mop_end; ++m, ++nm) {
        snprintf(buf, sizeof(buf), " %d", prog->h_mops[ps.mop_off + m].kind);
        kinds += buf;
      }
      kinds += "]";
    }
    snprintf(buf, sizeof(buf), "sweeps=%d mops=%d spawns=%d direct=%d/%d fast=%d", ps.nsweeps, nm,
             (int)ps.spawn_param.size(), ps.io_first & 1, (ps.io_last >> 1) & 1,
             ps.fast_ok ? 1 : (ps.seq_ok ? 2 : 0));
    return std::string(buf) + kinds;
  };
  if (prog->v1_grad_ok) {
    for (const V1Stage& sg : prog->v1_grad) {
      if (sg.type == 0) {
        snprintf(buf, sizeof(buf), "PASS partners=%d ", (int)sg.partners.size());
        s += std::string(buf) + pass_line(prog->v1_passes[sg.pass]) + "\n";
      } else if (sg.type == 1) {
        snprintf(buf, sizeof(buf), "GATHER params=%d\n", (int)sg.gather_params.size());
        s += buf;
      } else {
        snprintf(buf, sizeof(buf), "DOTS partners=%d\n", (int)sg.partners.size());
        s += buf;
      }
    }
  } else if (prog->v1_ok) {
    for (int pi : prog->v1_run) s += "PASS " + pass_line(prog->v1_passes[pi]) + "\n";
  } else {
    snprintf(buf, sizeof(buf), "v0 plan: %d run passes\n", (int)prog->run_passes.size());
    s += buf;
  }
  if (prog->bi_cut >= 0) {
    auto count = [](const pqc_program* q, int type) {
      int k = 0;
      for (const V1Stage& sg : q->v1_grad) k += sg.type == type;
      return k;
    };
    snprintf(buf, sizeof(buf),
             "BIDIR cut=%d/%d PF=%d PB=%d vector-passes %lld -> %lld | F passes=%d gathers=%d | "
             "M passes=%d | B passes=%d gathers=%d\n",
             prog->bi_cut, (int)prog->ops.size(), prog->bi_PF, prog->bi_PB, prog->fwd_cost,
             prog->bi_cost, count(prog->bi_F, 0), count(prog->bi_F, 1),
             (int)prog->bi_M->v1_run.size(), count(prog->bi_B, 0), count(prog->bi_B, 1));
    s += buf;
    // the three sub-programs of the meet-in-the-middle plan
    const pqc_program* subs[3] = {prog->bi_F, prog->bi_M, prog->bi_B};
    const char* names[3] = {"F", "M", "B"};
    for (int k = 0; k < 3; ++k) {
      const pqc_program* q = subs[k];
      auto sub_line = [&](const V1Pass& ps) {
        char b2[160];
        snprintf(b2, sizeof(b2), "  %s PASS sweeps=%d spawns=%d direct=%d/%d fast=%d\n", names[k],
                 ps.nsweeps, (int)ps.spawn_param.size(), ps.io_first & 1, (ps.io_last >> 1) & 1,
                 ps.fast_ok ? 1 : (ps.seq_ok ? 2 : 0));
        s += b2;
      };
      if (k == 1) {
        for (int pi : q->v1_run) sub_line(q->v1_passes[pi]);
      } else {
        for (const V1Stage& sg : q->v1_grad)
          if (sg.type == 0) sub_line(q->v1_passes[sg.pass]);
          else if (sg.type == 1) s += std::string("  ") + names[k] + " GATHER\n";
      }
    }
  }
  // the plain run plan (PQC.run), when it differs from what was printed above
  if (prog->v1_ok && prog->v1_grad_ok)
    for (int pi : prog->v1_run) s += "RUN PASS " + pass_line(prog->v1_passes[pi]) + "\n";
  if (prog->front_ok) {
    snprintf(buf, sizeof(buf), "FRONT plan: %d passes, used for PQC.run: %d\n", (int)prog->front_run.size(),
             pqc_use_front(prog) ? 1 : 0);
    s += buf;
    for (int pi : prog->front_run) s += "FRONT PASS " + pass_line(prog->v1_passes[pi]) + "\n";
  }
  if ((int64_t)s.size() + 1 > cap) s.resize((size_t)cap - 1);
  memcpy(out, s.c_str(), s.size() + 1);
  return 0;
}

// =====================================================================================
// sweep kernel
// =====================================================================================
// Per-sample trig table of one plan: entry [sample][slot] for every TrigJob.  Computed once
// per batch so the pass kernel's prologue is a plain 16-byte load per entry (no sincos, no
// dependent angle loads in front of the first barrier).
__global__ void __launch_bounds__(256) k_trig_fill(const TrigJob* __restrict__ jobs, int njobs,
                                                   const double* __restrict__ angles, long long ld,
                                                   long long S, int nslots,
                                                   double2* __restrict__ out) {
  const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
  if (e >= S * njobs) return;
  const long long sample = e / njobs;
  const TrigJob jb = jobs[e - sample * njobs];
  double2* t = out + sample * nslots + jb.slot;
  double th = jb.offset, s, c;
  if (jb.kind == PQC_OP_FSIM || jb.kind == PQC_OP_FIXED_FSIM) {
    if (jb.param >= 0) th += angles[sample * ld + jb.param];
    sincos(th, &s, &c);
    t[0] = make_double2(c, s);
    const double ph = jb.param2 >= 0 ? angles[sample * ld + jb.param2] : jb.scale;
    sincos(ph, &s, &c);
    t[1] = make_double2(c, s);
  } else {
    if (jb.param >= 0) th += jb.scale * angles[sample * ld + jb.param];
    if (jb.kind == PQC_K_ZZSUM) {
      // entry[k] = exp(-i th/2 (npairs - 2k)), k = jb.pad = number of anti-aligned pairs
      sincos(-0.5 * th * (double)(jb.npairs - 2 * jb.pad), &s, &c);
      t[0] = make_double2(c, s);
    } else if (jb.kind == PQC_K_RXY) {
      sincos(th, &s, &c);                      // rx-like rotation by the FULL angle
      t[0] = make_double2(c, s);
    } else if (jb.pad == 1) {                  // member of a 1-qubit layer op: (tan, cos)
      sincos(0.5 * th, &s, &c);
      t[0] = make_double2(s / c, c);
    } else {
      sincos(0.5 * th, &s, &c);
      t[0] = make_double2(c, s);
    }
  }
}

// grow-only scratch for the trig table of `prog` and one fill launch
static int trig_prepare(const pqc_program* cprog, int tj0, int ntj, int nslots,
                        const double* d_angles, long long ld, long long S, cudaStream_t st) {
  pqc_program* prog = const_cast<pqc_program*>(cprog);
  const size_t need = (size_t)std::max<long long>(1, S) * std::max(1, nslots);
  if (prog->trig_cap < need) {
    if (prog->d_trig) {
      PQC_CUDA(cudaStreamSynchronize(st));
      PQC_CUDA(cudaFree(prog->d_trig));
      prog->d_trig = nullptr;
      prog->trig_cap = 0;
    }
    PQC_CUDA(cudaMalloc(&prog->d_trig, need * sizeof(double2)));
    prog->trig_cap = need;
  }
  const long long tot = S * ntj;
  if (tot <= 0) return 0;
  if ((tot + 255) / 256 > 0x7fffffffLL) PQC_FAIL(-1, "trig grid too large; split the batch");
  k_trig_fill<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(prog->d_tjobs + tj0, ntj, d_angles, ld,
                                                             S, nslots, prog->d_trig);
  PQC_LAUNCH_CHECK();
  return 0;
}

struct V1Args {
  const c128* src;           // buffer read by normal items (ping) -- may equal dst (in place)
  c128* dst;                 // buffer written (pong)
  const MOp* mops;
  const SweepD* sweeps;
  int nsweeps;
  const double2* gtrig;      // per-sample trig table [S][tstride], filled by k_trig_fill
  int tstride, toff, ntrig;  // this pass' entries: gtrig[sample * tstride + toff + 0..ntrig)
  const uint32_t* wtab;      // this pass' linear-form tables (nwt x V1_WTAB words)
  int nwt;
  const GenTerm* gens;
  int n, tb, items_log2, low_run;
  int lbit[V1_LOCAL_BITS];
  int obit[PQC_MAX_QUBITS];
  long long n_items;         // samples * (active + nspawn)
  int slots_total, active, nspawn;
  int spawn_slot[V1_MAX_SPAWN], spawn_goff[V1_MAX_SPAWN], spawn_gcnt[V1_MAX_SPAWN];
  int npartners;
  int partner_slot[V1_MAX_PART];
  c128* gpart;               // [S][P+1][P][ntiles]
  int P, ntiles;
  int sweeps_nmops, sweep0_io, last_io;
  int pf_dist;               // L2 prefetch distance in CTAs (0 = off)
  const V1Pass* hpass;       // host side only: the pass (fast-path plan) and its program
  const pqc_program* hprog;
};


// initial state into slot 0 of every sample (mode 1 |0..0>, 2 broadcast, 3 per sample)
__global__ void __launch_bounds__(256) k_init_slot0(c128* __restrict__ buf, int mode,
                                                    const c128* __restrict__ init,
                                                    long long init_stride, long long S,
                                                    int slots_total, int n) {
  const long long D = 1ll << n, total = S * D;
  for (long long g = (long long)blockIdx.x * 256 + threadIdx.x; g < total;
       g += (long long)gridDim.x * 256) {
    const long long s = g >> n, x = g & (D - 1);
    c128 v;
    if (mode == 1) v = make_double2(x == 0 ? 1.0 : 0.0, 0.0);
    else if (mode == 2) v = init[x];
    else v = init[s * init_stride + x];
    buf[((s * slots_total) << n) + x] = v;
  }
}

#define V1_MAX_MOPS 160
#define V1_MAX_SWEEPS 32

// DOTS: take Gram partials against `partner_slot[]` while loading; GEN: the pass carries
// in-pass diagonal-generator spawn items.
template <bool DOTS, bool GEN>
__global__ void __launch_bounds__(V1_NT, 512 / V1_NT) k_sweep_pass(const V1Args A) {
  extern __shared__ __align__(16) unsigned char smraw[];
  c128* sm = reinterpret_cast<c128*>(smraw);
  double2* trig = reinterpret_cast<double2*>(sm + (1u << V1_LOCAL_BITS));
  __shared__ double red[32];
  __shared__ MOp s_mops[V1_MAX_MOPS];
  __shared__ SweepD s_sweeps[V1_MAX_SWEEPS];
  __shared__ uint32_t s_wt[V1_MAX_WT * V1_WTAB];
  const int tid = threadIdx.x;
  const int tiles_log2 = A.n - A.tb;
  const long long blk = blockIdx.x;
  const long long group = blk >> tiles_log2;
  const uint32_t tile = (uint32_t)(blk & ((1ll << tiles_log2) - 1));
  uint32_t tbase = 0;
  for (int j = 0; j < tiles_log2; ++j) tbase |= ((tile >> j) & 1u) << A.obit[j];
  const int ipc = 1 << A.items_log2;
  const long long item0 = group << A.items_log2;
  const int ips = A.active + A.nspawn;              // items per sample
  const uint32_t amask = (1u << A.tb) - 1u;

  const long long sample0 = item0 / ips;
  const int r0 = (int)(item0 - sample0 * ips);
  // item li of this CTA (li < 16): 32-bit arithmetic only
  auto item_info = [&](int li, long long& sample, int& src_slot, int& dst_slot, int& gen) {
    const int rr = r0 + li, ds = rr / ips, r = rr - ds * ips;
    sample = sample0 + ds;
    if (!GEN || r < A.active) {
      src_slot = dst_slot = r;
      gen = -1;
    } else {
      src_slot = 0;
      dst_slot = A.spawn_slot[r - A.active];
      gen = r - A.active;
    }
  };
  auto local_to_amp = [&](uint32_t i) -> uint32_t {
    uint32_t r = i & ((1u << A.low_run) - 1u);
    for (int j = A.low_run; j < A.tb; ++j) r |= ((i >> j) & 1u) << A.lbit[j];
    return r;
  };
  // staged (global <-> swizzled shared) element r of this thread: local index tid | r << V1_TBITS
  const uint32_t st_amp_tid = tbase | local_to_amp((uint32_t)tid & amask);
  uint32_t st_h[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) st_h[k] = (V1_TBITS + k < A.tb) ? (1u << A.lbit[V1_TBITS + k]) : 0u;
  const uint32_t st_s_tid = swz((uint32_t)tid);
  const uint32_t st_s[4] = {swz(1u << V1_TBITS), swz(2u << V1_TBITS), swz(4u << V1_TBITS),
                            swz(8u << V1_TBITS)};
  const int st_item_shift = A.tb - V1_TBITS;    // li = r >> (tb - V1_TBITS)   (tb >= 8)

  const bool direct_load = (A.sweep0_io & 1) != 0;
  long long my_sample = 0;
  int my_src = 0, my_dst = 0, my_gen = -1;
  if (ipc == 1) item_info(0, my_sample, my_src, my_dst, my_gen);

  // ---- early tile load: when sweep 0 reads global memory directly (and has no index
  // permutation in front), issue its 16 loads NOW so their latency overlaps the prologue
  c128 a[16];
  bool preloaded = false;
  if (!DOTS && direct_load && ipc == 1) {     // (the in-pass Gram path loads + dots together)
    const SweepD* sw = A.sweeps;
    if ((sw->pad & 0xffff) == 0) {
      const uint32_t amp0 = tbase | __ldg(&sw->tamp[0][tid & 15]) | __ldg(&sw->tamp[1][tid >> 4]);
      const uint32_t g0 = sw->gm[0], g1 = sw->gm[1], g2 = sw->gm[2], g3 = sw->gm[3];
      const c128* sp = A.src + ((my_sample * A.slots_total + my_src) << A.n) + amp0;
#pragma unroll
      for (int j = 0; j < 16; ++j) a[j] = sp[XSEL4R(j, g0, g1, g2, g3)];
      preloaded = true;
    }
  }

  // ---- L2 prefetch of the tile a later CTA (blockIdx + pf_dist, about one wave ahead) will
  // load: one 256-byte run per thread, fire and forget.  Turns that CTA's HBM latency into an
  // L2 hit; the data waits in the 126 MB L2 for roughly one CTA lifetime.
  if (A.pf_dist > 0 && ipc == 1 && A.tb == V1_LOCAL_BITS) {
    const long long blk2 = blk + A.pf_dist;
    if (blk2 < (long long)gridDim.x) {
      const long long item2 = blk2 >> tiles_log2;
      const uint32_t tile2 = (uint32_t)(blk2 & ((1ll << tiles_log2) - 1));
      uint32_t amp2 = 0;
      for (int j = 0; j < tiles_log2; ++j) amp2 |= ((tile2 >> j) & 1u) << A.obit[j];
      amp2 |= local_to_amp((uint32_t)tid << 4);
      const long long sample2 = item2 / ips;
      const int r2 = (int)(item2 - sample2 * ips);
      const int slot2 = (!GEN || r2 < A.active) ? r2 : 0;
      const c128* pa = A.src + ((sample2 * A.slots_total + slot2) << A.n) + amp2;
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], 256;" ::"l"(pa) : "memory");
    }
  }

  // ---- stage the micro-program and fill the per-item trig table ------------------------------
  {
    const int nm = A.sweeps_nmops;
    int* d = reinterpret_cast<int*>(s_mops);
    const int* g = reinterpret_cast<const int*>(A.mops);
    for (int e = tid; e < nm * (int)(sizeof(MOp) / 4); e += V1_NT) d[e] = g[e];
    int* d2 = reinterpret_cast<int*>(s_sweeps);
    const int* g2 = reinterpret_cast<const int*>(A.sweeps);
    for (int e = tid; e < A.nsweeps * (int)(sizeof(SweepD) / 4); e += V1_NT) d2[e] = g2[e];
  }
  for (int e = tid; e < A.nwt * V1_WTAB; e += V1_NT) s_wt[e] = A.wtab[e];
  for (int e = tid; e < ipc * A.ntrig; e += V1_NT) {
    const int li = e / A.ntrig;
    const long long item = item0 + li;
    if (item >= A.n_items) continue;
    trig[e] = A.gtrig[(item / ips) * A.tstride + A.toff + (e - li * A.ntrig)];
  }

  // ---- staged load (global -> swizzled shared) when the first sweep cannot load directly -------
  if (!direct_load) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int li = r >> st_item_shift;
      c128 v = make_double2(0.0, 0.0);
      if (item0 + li < A.n_items) {
        long long sample = my_sample;
        int ss = my_src, ds = my_dst, gg = my_gen;
        if (ipc > 1) item_info(li, sample, ss, ds, gg);
        v = A.src[((sample * A.slots_total + ss) << A.n) +
                  (st_amp_tid | SEL4R(r, st_h[0], st_h[1], st_h[2], st_h[3]))];
      }
      sm[st_s_tid ^ XSEL4R(r, st_s[0], st_s[1], st_s[2], st_s[3])] = v;
    }
    if (DOTS && my_gen < 0) {
      const c128* own = A.src + ((my_sample * A.slots_total + my_src) << A.n);
      for (int q = 0; q < A.npartners; ++q) {
        const int ps = A.partner_slot[q];
        if (my_src > ps) continue;
        const c128* pv = A.src + ((my_sample * A.slots_total + ps) << A.n);
        double re = 0.0, im = 0.0;
        for (uint32_t i = tid; i < (1u << V1_LOCAL_BITS); i += V1_NT) {
          const uint32_t amp = tbase | local_to_amp(i);
          const c128 x = own[amp], y = pv[amp];
          re += x.x * y.x + x.y * y.y;
          im += x.x * y.y - x.y * y.x;
        }
        re = block_sum<V1_NT>(re, red);
        im = block_sum<V1_NT>(im, red);
        if (tid == 0)
          A.gpart[((my_sample * (A.P + 1) + my_src) * A.P + (ps - 1)) * A.ntiles + tile] =
              make_double2(re, im);
      }
    }
  }
  __syncthreads();

  // ---- sweeps ------------------------------------------------------------------------------------
  // product of the cos factors of the tangent-form layer ops; a scalar, so with one item per
  // CTA it is carried across sweeps and multiplied in once before the final store
  double fscale = 1.0;
  for (int si = 0; si < A.nsweeps; ++si) {
    const SweepD& sw = s_sweeps[si];
    // the 8 thread bits go to the non-register local positions (tables from the planner)
#ifdef V1_ALU_PROLOGUE
    uint32_t base = 0, amp0 = tbase;
#pragma unroll
    for (int t = 0; t < V1_TBITS; ++t) {
      const uint32_t bit = ((uint32_t)tid >> t) & 1u;
      base |= bit << sw.tl[t];
      amp0 |= sw.tg[t] < 32 ? (bit << sw.tg[t]) : 0u;
    }
#else
    const uint32_t base = (uint32_t)sw.tpos[0][tid & 15] | (uint32_t)sw.tpos[1][tid >> 4];
    const uint32_t amp0 = tbase | sw.tamp[0][tid & 15] | sw.tamp[1][tid >> 4];
#endif
    // swz is linear over GF(2): swz(base | sel) = swz(base) ^ swz(sel)
    const uint32_t sb = swz(base), s0 = sw.sz[0], s1 = sw.sz[1], s2 = sw.sz[2], s3 = sw.sz[3];
    // global amplitude index of register j = amp0 | (selected g-masks)
    const uint32_t g0 = sw.gm[0], g1 = sw.gm[1], g2 = sw.gm[2], g3 = sw.gm[3];
    // X / CNOT are affine maps of the 4-bit register index: pi(j) = XOR_{k in j} col[k] ^ v.
    // They cost nothing per amplitude: they only change the load / store address constants.
    const int npre = sw.pad & 0xffff, npost = sw.pad >> 16;
    auto affine = [&](int begin, int end, bool reverse, uint32_t (&col)[4], uint32_t& v) {
      col[0] = 1u; col[1] = 2u; col[2] = 4u; col[3] = 8u;
      v = 0u;
      for (int q = 0; q < end - begin; ++q) {
        const MOp& pm = s_mops[reverse ? end - 1 - q : begin + q];
        if (pm.kind == PQC_OP_X) {
          v ^= 1u << pm.k0;
        } else if (pm.k0 >= 0) {               // CNOT, control in registers
          const int kc = pm.k0, kt = pm.k1;
#pragma unroll
          for (int c = 0; c < 4; ++c) col[c] ^= ((col[c] >> kc) & 1u) << kt;
          v ^= ((v >> kc) & 1u) << kt;
        } else {                               // CNOT, control fixed for this thread
          const uint32_t cb = pm.l0 >= 0 ? ((base >> pm.l0) & 1u) : ((tbase >> pm.b0) & 1u);
          v ^= cb << pm.k1;
        }
      }
    };
#define LIN4(x, a0, a1, a2, a3) \
  ((((x)&1u) ? (a0) : 0u) ^ (((x)&2u) ? (a1) : 0u) ^ (((x)&4u) ? (a2) : 0u) ^ (((x)&8u) ? (a3) : 0u))
    // w(amp0) and the words of the 4 register bits for a linear-form table (ZZSUM / GEN)
    auto lin_words = [&](const uint32_t* wt, const SweepD& d, uint32_t& w0, uint32_t& w1,
                         uint32_t& w2, uint32_t& w3, uint32_t& w4) {
      uint32_t w = 0u;
      for (int jb = 0; jb < tiles_log2; ++jb) w ^= ((tile >> jb) & 1u) ? wt[A.obit[jb]] : 0u;
#pragma unroll
      for (int t = 0; t < V1_TBITS; ++t) {
        const int gbit = min((int)d.tg[t], 32);
        w ^= ((tid >> t) & 1) ? wt[gbit] : 0u;
      }
      w0 = w;
      w1 = wt[d.gb[0]]; w2 = wt[d.gb[1]]; w3 = wt[d.gb[2]]; w4 = wt[d.gb[3]];
    };
    const int li = (int)(base >> A.tb);
    long long sample = my_sample;
    int src_slot = my_src, dst_slot = my_dst, gen = my_gen;
    const bool live = item0 + li < A.n_items;
    if (ipc > 1 && live) item_info(li, sample, src_slot, dst_slot, gen);
    const double2* tg = trig + (size_t)li * A.ntrig;
#define SEL4(j, a0, a1, a2, a3) \
  ((((j)&1) ? (a0) : 0u) | (((j)&2) ? (a1) : 0u) | (((j)&4) ? (a2) : 0u) | (((j)&8) ? (a3) : 0u))
#define XSEL4(j, a0, a1, a2, a3) \
  ((((j)&1) ? (a0) : 0u) ^ (((j)&2) ? (a1) : 0u) ^ (((j)&4) ? (a2) : 0u) ^ (((j)&8) ? (a3) : 0u))

    const bool dl = si == 0 && direct_load;
    uint32_t lc[4] = {1u, 2u, 4u, 8u}, lv = 0u;
    if (npre) affine(sw.mop_begin, sw.mop_begin + npre, true, lc, lv);
    if (dl && preloaded) {
      // loaded before the prologue
    } else if (dl) {
      uint32_t lgb = amp0, lg0 = g0, lg1 = g1, lg2 = g2, lg3 = g3;
      if (npre) {
        lgb = amp0 ^ LIN4(lv, g0, g1, g2, g3);
        lg0 = LIN4(lc[0], g0, g1, g2, g3); lg1 = LIN4(lc[1], g0, g1, g2, g3);
        lg2 = LIN4(lc[2], g0, g1, g2, g3); lg3 = LIN4(lc[3], g0, g1, g2, g3);
      }
      const c128* sp = A.src + ((sample * A.slots_total + src_slot) << A.n);
#pragma unroll
      for (int j = 0; j < 16; ++j)
        a[j] = live ? sp[lgb ^ XSEL4(j, lg0, lg1, lg2, lg3)] : make_double2(0.0, 0.0);
      if (DOTS && gen < 0) {
        // Gram partials against every pending partner: per-warp shuffles, then ONE barrier
        __shared__ double dred[V1_NT / 32][2 * V1_MAX_PART];
        const int w = tid >> 5, l = tid & 31;
        for (int q = 0; q < A.npartners; ++q) {
          const int ps = A.partner_slot[q];
          if (src_slot > ps) continue;               // uniform per CTA
          const c128* pv = A.src + ((sample * A.slots_total + ps) << A.n);
          double re = 0.0, im = 0.0;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const c128 y = pv[lgb ^ XSEL4(j, lg0, lg1, lg2, lg3)];
            re += a[j].x * y.x + a[j].y * y.y;
            im += a[j].x * y.y - a[j].y * y.x;
          }
          re = warp_sum(re);
          im = warp_sum(im);
          if (l == 0) { dred[w][2 * q] = re; dred[w][2 * q + 1] = im; }
        }
        __syncthreads();
        if (tid < 2 * A.npartners) {
          const int q = tid >> 1, ps = A.partner_slot[q];
          if (src_slot <= ps) {
            double v = 0.0;
#pragma unroll
            for (int ww = 0; ww < V1_NT / 32; ++ww) v += dred[ww][tid];
            double* g = reinterpret_cast<double*>(
                A.gpart + ((sample * (A.P + 1) + src_slot) * A.P + (ps - 1)) * A.ntiles + tile);
            g[tid & 1] = v;
          }
        }
      }
    } else {
      uint32_t lsb = sb, ls0 = s0, ls1 = s1, ls2 = s2, ls3 = s3;
      if (npre) {
        lsb = sb ^ LIN4(lv, s0, s1, s2, s3);
        ls0 = LIN4(lc[0], s0, s1, s2, s3); ls1 = LIN4(lc[1], s0, s1, s2, s3);
        ls2 = LIN4(lc[2], s0, s1, s2, s3); ls3 = LIN4(lc[3], s0, s1, s2, s3);
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) a[j] = sm[lsb ^ XSEL4(j, ls0, ls1, ls2, ls3)];
    }

    for (int mi = sw.mop_begin + npre; mi < sw.mop_end - npost; ++mi) {
      const MOp& m = s_mops[mi];
#if defined(V1_ABL_NOOPS)           // timing ablations (tools/microbench.py); results are wrong
      const int kind = PQC_OP_IDENT; (void)m;
#elif defined(V1_ABL_NORX)
      const int kind = m.kind == PQC_K_LAYER_RX4 ? PQC_OP_IDENT : m.kind;
#elif defined(V1_ABL_NOZZ)
      const int kind = m.kind == PQC_K_ZZSUM ? PQC_OP_IDENT : m.kind;
#else
      const int kind = m.kind;
#endif
      if (kind == PQC_K_LAYER_RX4) {
        // straight-line over the 4 register bits; absent gates are the identity (t 0, c 1).
        // trig entries of layer ops hold (tan, cos) of the half angle.
        const int sk = m.subk;
        double2 c0 = make_double2(0.0, 1.0), c1 = c0, c2 = c0, c3 = c0;
        if (sk & 0xff) c0 = tg[m.subt[0]];
        if (sk & 0xff00) c1 = tg[m.subt[1]];
        if (sk & 0xff0000) c2 = tg[m.subt[2]];
        if (sk & 0xff000000) c3 = tg[m.subt[3]];
        op_rx_t<0>(a, c0.x);
        op_rx_t<1>(a, c1.x);
        op_rx_t<2>(a, c2.x);
        op_rx_t<3>(a, c3.x);
        fscale *= (c0.y * c1.y) * (c2.y * c3.y);
      } else if (kind == PQC_K_LAYER_REAL4) {
        const int sk = m.subk;
        double f = 1.0;
#define REAL_SLOT(K)                                                       \
  {                                                                        \
    const int kd = ((sk >> (8 * K)) & 0xff) - 1;                           \
    if (kd == PQC_OP_RY) { const double2 tc = tg[m.subt[K]]; op_ry_t<K>(a, tc.x); f *= tc.y; } \
    else if (kd == PQC_OP_H) { op_h_u<K>(a); f *= 0.70710678118654752440; } \
  }
        REAL_SLOT(0) REAL_SLOT(1) REAL_SLOT(2) REAL_SLOT(3)
#undef REAL_SLOT
        fscale *= f;
      } else if (kind == PQC_K_ZZSUM) {
        // phase = table[number of anti-aligned pairs of x] = table[popc(w(x))], w linear
        uint32_t w0, w1, w2, w3, w4;
        lin_words(s_wt + m.aux1, sw, w0, w1, w2, w3, w4);
        const double2* tz = tg + m.trig;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const double2 ph = tz[__popc(w0 ^ XSEL4(j, w1, w2, w3, w4))];
          const c128 v = a[j];
          a[j] = make_double2(v.x * ph.x - v.y * ph.y, v.y * ph.x + v.x * ph.y);
        }
      } else if (kind == PQC_OP_RZ || kind == PQC_OP_S || kind == PQC_OP_T) {
        double c = 1.0, s = 0.0;
        if (kind == PQC_OP_RZ) { const double2 cs = tg[m.trig]; c = cs.x; s = cs.y; }
        const int cb = m.l0 >= 0 ? (int)((base >> m.l0) & 1u) : (int)((tbase >> m.b0) & 1u);
        const int k0 = m.k0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int bit = k0 >= 0 ? ((j >> k0) & 1) : cb;
          c128 v = a[j];
          if (kind == PQC_OP_RZ) {
            const double sg = bit ? s : -s;
            v = make_double2(v.x * c - v.y * sg, v.y * c + v.x * sg);
          } else if (bit) {
            if (kind == PQC_OP_S) {
              v = make_double2(-v.y, v.x);
            } else {
              const double r = 0.70710678118654752440;
              v = make_double2(r * (v.x - v.y), r * (v.x + v.y));
            }
          }
          a[j] = v;
        }
      } else if (kind == PQC_OP_CZ || kind == PQC_OP_RZZ) {
        double c = 1.0, s = 0.0;
        if (kind == PQC_OP_RZZ) { const double2 cs = tg[m.trig]; c = cs.x; s = cs.y; }
        const int ca = m.l0 >= 0 ? (int)((base >> m.l0) & 1u) : (int)((tbase >> m.b0) & 1u);
        const int cb = m.l1 >= 0 ? (int)((base >> m.l1) & 1u) : (int)((tbase >> m.b1) & 1u);
        const int k0 = m.k0, k1 = m.k1;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int ba = k0 >= 0 ? ((j >> k0) & 1) : ca;
          const int bb = k1 >= 0 ? ((j >> k1) & 1) : cb;
          c128 v = a[j];
          if (kind == PQC_OP_CZ) {
            if (ba & bb) v = make_double2(-v.x, -v.y);
          } else {
            const double sg = (ba ^ bb) ? s : -s;
            v = make_double2(v.x * c - v.y * sg, v.y * c + v.x * sg);
          }
          a[j] = v;
        }
      } else if (kind == PQC_K_GEN) {
        if (GEN && gen == m.aux0) {
          // uniform-coefficient diagonal generator: c0 * sum_t (-1)^{popc(x & z_t)}
          //   = c0 * (nterms - 2 popc(w(x)))
          uint32_t w0, w1, w2, w3, w4;
          lin_words(s_wt + m.aux1, sw, w0, w1, w2, w3, w4);
          const double cr = A.gens[A.spawn_goff[gen]].re, ci = A.gens[A.spawn_goff[gen]].im;
          const int nt = m.npairs;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const double f = (double)(nt - 2 * __popc(w0 ^ XSEL4(j, w1, w2, w3, w4)));
            const double fr = f * cr, fi = f * ci;
            const c128 v = a[j];
            a[j] = make_double2(v.x * fr - v.y * fi, v.y * fr + v.x * fi);
          }
        }
      } else if (kind != PQC_OP_IDENT) {
        // symmetric two-bit rotations, both bits in registers
        double ce = 1.0, se = 0.0, co = 1.0, so = 0.0, pc = 1.0, psn = 0.0;
        bool even = false, ph = false;
        if (kind == PQC_OP_RXX) { const double2 cs = tg[m.trig]; ce = co = cs.x; se = so = cs.y; even = true; }
        else if (kind == PQC_OP_RYY) { const double2 cs = tg[m.trig]; ce = co = cs.x; so = cs.y; se = -cs.y; even = true; }
        else if (kind == PQC_K_RXY) { const double2 cs = tg[m.trig]; co = cs.x; so = cs.y; }
        else if (kind == PQC_OP_SQRTISWAP) { co = 0.70710678118654752440; so = -co; }
        else { const double2 cs = tg[m.trig]; co = cs.x; so = cs.y;
               if (kind == PQC_OP_FSIM) { const double2 p2 = tg[m.trig + 1]; pc = p2.x; psn = p2.y; ph = true; } }
        const int ka = m.k0 < m.k1 ? m.k0 : m.k1, kb = m.k0 < m.k1 ? m.k1 : m.k0;
        switch (ka * 4 + kb) {
          case 1: op_pair<0, 1>(a, ce, se, co, so, even, ph, pc, psn); break;
          case 2: op_pair<0, 2>(a, ce, se, co, so, even, ph, pc, psn); break;
          case 3: op_pair<0, 3>(a, ce, se, co, so, even, ph, pc, psn); break;
          case 6: op_pair<1, 2>(a, ce, se, co, so, even, ph, pc, psn); break;
          case 7: op_pair<1, 3>(a, ce, se, co, so, even, ph, pc, psn); break;
          default: op_pair<2, 3>(a, ce, se, co, so, even, ph, pc, psn); break;
        }
      }
    }

#ifdef V1_SCALE_PER_SWEEP
    if (fscale != 1.0) {
#else
    if (fscale != 1.0 && (ipc > 1 || si + 1 == A.nsweeps)) {   // uniform per item
#endif
      op_scale(a, fscale);
      fscale = 1.0;
    }
    const bool ds = (si + 1 == A.nsweeps) && (sw.io & 2);
    uint32_t sc[4] = {1u, 2u, 4u, 8u}, sv = 0u;
    if (npost) affine(sw.mop_end - npost, sw.mop_end, false, sc, sv);
    if (ds) {
      uint32_t sgb = amp0, sg0 = g0, sg1 = g1, sg2 = g2, sg3 = g3;
      if (npost) {
        sgb = amp0 ^ LIN4(sv, g0, g1, g2, g3);
        sg0 = LIN4(sc[0], g0, g1, g2, g3); sg1 = LIN4(sc[1], g0, g1, g2, g3);
        sg2 = LIN4(sc[2], g0, g1, g2, g3); sg3 = LIN4(sc[3], g0, g1, g2, g3);
      }
      if (live) {
        c128* dp = A.dst + ((sample * A.slots_total + dst_slot) << A.n);
#pragma unroll
        for (int j = 0; j < 16; ++j) dp[sgb ^ XSEL4(j, sg0, sg1, sg2, sg3)] = a[j];
      }
    } else {
      uint32_t ssb = sb, ss0 = s0, ss1 = s1, ss2 = s2, ss3 = s3;
      if (npost) {
        ssb = sb ^ LIN4(sv, s0, s1, s2, s3);
        ss0 = LIN4(sc[0], s0, s1, s2, s3); ss1 = LIN4(sc[1], s0, s1, s2, s3);
        ss2 = LIN4(sc[2], s0, s1, s2, s3); ss3 = LIN4(sc[3], s0, s1, s2, s3);
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) sm[ssb ^ XSEL4(j, ss0, ss1, ss2, ss3)] = a[j];
      __syncthreads();
    }
  }

  // ---- staged store -------------------------------------------------------------------------------
  if (!(A.last_io & 2)) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int li = r >> st_item_shift;
      if (item0 + li < A.n_items) {
        long long sample = my_sample;
        int ss = my_src, ds = my_dst, gg = my_gen;
        if (ipc > 1) item_info(li, sample, ss, ds, gg);
        A.dst[((sample * A.slots_total + ds) << A.n) +
              (st_amp_tid | SEL4R(r, st_h[0], st_h[1], st_h[2], st_h[3]))] =
            sm[st_s_tid ^ XSEL4R(r, st_s[0], st_s[1], st_s[2], st_s[3])];
      }
    }
  }
}

// =====================================================================================
// layer pass: the fast path for passes built from the three aligned nibble sweeps
// (FastPlan, pqc_common.cuh).  Same arithmetic, same shared-memory swizzle and the same op
// order as k_sweep_pass, but the pass description sits in the kernel arguments and the sweep
// geometry is compile-time: shared-memory addresses are `base ^ constant`, the thread ->
// amplitude maps are two 16-entry tables and the linear forms of the ZZSUM / GEN ops are
// pre-folded per nibble value, so the per-amplitude integer work is one LOP3 + POPC.
// =====================================================================================
#ifdef FAST_STREAM_HINTS     // A/B: evict-first hints on the streaming tile traffic
#define LP_LD(p) __ldcs(p)
#define LP_ST(p, v) __stcs(p, v)
#else
#define LP_LD(p) (*(p))
#define LP_ST(p, v) (*(p) = (v))
#endif
struct FastArgs {
  const c128* src;
  c128* dst;
  const double2* gtrig;
  int tstride, toff, ntrig;
  const uint32_t* wtab;
  int nwt;
  int n;
  int lbit[12];
  int obit[PQC_MAX_QUBITS];
  int slots_total, active, nspawn;
  int spawn_slot[FAST_MAX_SPAWN];
  double spawn_cr[FAST_MAX_SPAWN], spawn_ci[FAST_MAX_SPAWN];
  int pf_dist;               // L2 prefetch distance in CTAs (0 = off)
  FastPlan plan;
};

// the pass' ops on the 16 register amplitudes; RN = tile nibble held in registers, (NX, vx) and
// (NY, vy) = the two nibbles fed by the thread index and this thread's values for them
template <bool GEN, int RN, int NX, int NY>
__device__ __forceinline__ void fast_ops(c128 (&a)[16], const FastArgs& A, int slot, int vx, int vy,
                                         const double2* trig,
                                         const uint32_t (*s_wn)[3][16], const uint32_t* s_wb,
                                         int gen, double& fscale) {
  const int nops = A.plan.nops[slot];
  for (int oi = 0; oi < nops; ++oi) {
    const FastOp& op = A.plan.ops[slot][oi];
    const int kind = op.kind;
    if (kind == PQC_K_LAYER_RX4) {
      const int sk = op.subk;
      double2 c0 = make_double2(0.0, 1.0), c1 = c0, c2 = c0, c3 = c0;
      if (sk & 0xff) c0 = trig[op.t[0]];
      if (sk & 0xff00) c1 = trig[op.t[1]];
      if (sk & 0xff0000) c2 = trig[op.t[2]];
      if (sk & 0xff000000) c3 = trig[op.t[3]];
      op_rx_t<0>(a, c0.x);
      op_rx_t<1>(a, c1.x);
      op_rx_t<2>(a, c2.x);
      op_rx_t<3>(a, c3.x);
      fscale *= (c0.y * c1.y) * (c2.y * c3.y);
    } else if (kind == PQC_K_LAYER_REAL4) {
      const int sk = op.subk;
      double f = 1.0;
#define FAST_REAL_SLOT(K)                                                  \
  {                                                                        \
    const int kd = ((sk >> (8 * K)) & 0xff) - 1;                           \
    if (kd == PQC_OP_RY) { const double2 tc = trig[op.t[K]]; op_ry_t<K>(a, tc.x); f *= tc.y; } \
    else if (kd == PQC_OP_H) { op_h_u<K>(a); f *= 0.70710678118654752440; } \
  }
      FAST_REAL_SLOT(0) FAST_REAL_SLOT(1) FAST_REAL_SLOT(2) FAST_REAL_SLOT(3)
#undef FAST_REAL_SLOT
      fscale *= f;
    } else {
      // ZZSUM / GEN: w(x) = w(tile) ^ w(thread nibbles) ^ w(register nibble value j)
      const uint32_t(*wn)[16] = s_wn[op.wt];
      const uint32_t w0 = s_wb[op.wt] ^ wn[NX][vx] ^ wn[NY][vy];
      const uint32_t w1 = wn[RN][1], w2 = wn[RN][2], w3 = wn[RN][4], w4 = wn[RN][8];
      if (kind == PQC_K_ZZSUM) {
        const double2* tz = trig + op.t[0];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const double2 ph = tz[__popc(w0 ^ XSEL4R(j, w1, w2, w3, w4))];
          const c128 v = a[j];
          a[j] = make_double2(v.x * ph.x - v.y * ph.y, v.y * ph.x + v.x * ph.y);
        }
      } else if (GEN && gen == op.spawn) {
        const double cr = A.spawn_cr[gen], ci = A.spawn_ci[gen];
        const int nt = op.nterms;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const double f = (double)(nt - 2 * __popc(w0 ^ XSEL4R(j, w1, w2, w3, w4)));
          const double fr = f * cr, fi = f * ci;
          const c128 v = a[j];
          a[j] = make_double2(v.x * fr - v.y * fi, v.y * fr + v.x * fi);
        }
      }
    }
  }
}

template <int NS, bool GEN>
__global__ void __launch_bounds__(256, 2) k_layer_pass(const FastArgs A) {
  extern __shared__ __align__(16) c128 lp_sm[];
  double2* trig = reinterpret_cast<double2*>(lp_sm + 4096);
  __shared__ uint32_t s_ta[2][16];               // amplitude bits of tile nibbles 1 and 2
  __shared__ uint32_t s_wn[FAST_MAX_WT][3][16];  // linear forms per table, nibble, nibble value
  __shared__ uint32_t s_wb[FAST_MAX_WT];         // linear form of the tile index
  const int tid = threadIdx.x, lo = tid & 15, hi = tid >> 4;
  const int tiles_log2 = A.n - 12;
  const long long item = (long long)blockIdx.x >> tiles_log2;
  const uint32_t tile = (uint32_t)(blockIdx.x & ((1u << tiles_log2) - 1u));
  uint32_t tbase = 0;
  for (int j = 0; j < tiles_log2; ++j) tbase |= ((tile >> j) & 1u) << A.obit[j];
  const int ips = A.active + A.nspawn;
  const long long sample = item / ips;
  const int r = (int)(item - sample * ips);
  int src_slot = r, dst_slot = r, gen = -1;
  if (GEN && r >= A.active) {
    src_slot = 0;
    gen = r - A.active;
    dst_slot = A.spawn_slot[gen];
  }
  // ---- sweep A loads straight from global memory: registers = tile positions 8-11
  c128 a[16];
  {
    const uint32_t ampA = tbase | (uint32_t)lo | ((uint32_t)(hi & 1) << A.lbit[4]) |
                          ((uint32_t)((hi >> 1) & 1) << A.lbit[5]) |
                          ((uint32_t)((hi >> 2) & 1) << A.lbit[6]) |
                          ((uint32_t)((hi >> 3) & 1) << A.lbit[7]);
    const uint32_t g0 = 1u << A.lbit[8], g1 = 1u << A.lbit[9], g2 = 1u << A.lbit[10],
                   g3 = 1u << A.lbit[11];
    const c128* sp = A.src + ((sample * A.slots_total + src_slot) << A.n) + ampA;
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = LP_LD(sp + XSEL4R(j, g0, g1, g2, g3));
  }
  // ---- L2 prefetch of the tile CTA blockIdx + pf_dist will load (one 256-byte run per thread)
  if (A.pf_dist > 0) {
    const long long blk2 = (long long)blockIdx.x + A.pf_dist;
    if (blk2 < (long long)gridDim.x) {
      const long long item2 = blk2 >> tiles_log2;
      const uint32_t tile2 = (uint32_t)(blk2 & ((1ll << tiles_log2) - 1));
      uint32_t amp2 = 0;
      for (int j = 0; j < tiles_log2; ++j) amp2 |= ((tile2 >> j) & 1u) << A.obit[j];
#pragma unroll
      for (int i = 0; i < 8; ++i) amp2 |= (((uint32_t)tid >> i) & 1u) << A.lbit[4 + i];
      const long long sample2 = item2 / ips;
      const int r2 = (int)(item2 - sample2 * ips);
      const int slot2 = (!GEN || r2 < A.active) ? r2 : 0;
      const c128* pa = A.src + ((sample2 * A.slots_total + slot2) << A.n) + amp2;
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], 256;" ::"l"(pa) : "memory");
    }
  }
  // ---- tables (overlap the tile load)
  if (tid < 32) {
    uint32_t v = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) v |= (((uint32_t)lo >> i) & 1u) << A.lbit[4 + 4 * hi + i];
    s_ta[hi][lo] = v;
  }
  for (int e = tid; e < A.nwt * 48; e += 256) (&s_wn[0][0][0])[e] = (&A.plan.wn[0][0][0])[e];
  if (tid < A.nwt) {
    uint32_t w = 0;
    for (int j = 0; j < tiles_log2; ++j) w ^= ((tile >> j) & 1u) ? A.plan.wo[tid][j] : 0u;
    s_wb[tid] = w;
  }
  for (int e = tid; e < A.ntrig; e += 256) trig[e] = A.gtrig[sample * A.tstride + A.toff + e];
  __syncthreads();

  double fscale = 1.0;
  // swizzled shared-memory index of register j: swz(base) ^ swz(j << shift)
#define LP_CA(j) ((((j) << 8) ^ ((((j) << 2) ^ ((j) >> 1)) & 7)))   /* swz(j << 8) */
#define LP_CB(j) (((j) ^ ((j) >> 3)))                               /* swz(j)      */
#define LP_CC(j) ((((j) << 4) ^ ((((j) << 1) ^ ((j) >> 2)) & 7)))   /* swz(j << 4) */
  fast_ops<GEN, 2, 0, 1>(a, A, 0, lo, hi, trig, s_wn, s_wb, gen, fscale);
  if (NS == 1) {
    // single sweep: store the registers where they came from
    if (fscale != 1.0) op_scale(a, fscale);
    const uint32_t ampA = tbase | (uint32_t)lo | s_ta[0][hi];
    const uint32_t g0 = 1u << A.lbit[8], g1 = 1u << A.lbit[9], g2 = 1u << A.lbit[10],
                   g3 = 1u << A.lbit[11];
    c128* dp = A.dst + ((sample * A.slots_total + dst_slot) << A.n) + ampA;
#pragma unroll
    for (int j = 0; j < 16; ++j) LP_ST(dp + XSEL4R(j, g0, g1, g2, g3), a[j]);
    return;
  }
  {
    const uint32_t sb = swz((uint32_t)tid);
#pragma unroll
    for (int j = 0; j < 16; ++j) lp_sm[sb ^ LP_CA(j)] = a[j];
  }
  __syncthreads();
  if (NS == 3) {
    // sweep B: registers = positions 0-3, threads = positions 4-11
    const uint32_t sb = swz((uint32_t)tid << 4);
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = lp_sm[sb ^ LP_CB(j)];
    fast_ops<GEN, 0, 1, 2>(a, A, 1, lo, hi, trig, s_wn, s_wb, gen, fscale);
#pragma unroll
    for (int j = 0; j < 16; ++j) lp_sm[sb ^ LP_CB(j)] = a[j];
    __syncthreads();
  }
  {
    // sweep C: registers = positions 4-7, threads = positions 0-3 and 8-11; stores to global
    const uint32_t sb = swz((uint32_t)lo | ((uint32_t)hi << 8));
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = lp_sm[sb ^ LP_CC(j)];
    fast_ops<GEN, 1, 0, 2>(a, A, 2, lo, hi, trig, s_wn, s_wb, gen, fscale);
    if (fscale != 1.0) op_scale(a, fscale);
    const uint32_t ampC = tbase | (uint32_t)lo | s_ta[1][hi];
    const uint32_t g0 = 1u << A.lbit[4], g1 = 1u << A.lbit[5], g2 = 1u << A.lbit[6],
                   g3 = 1u << A.lbit[7];
    c128* dp = A.dst + ((sample * A.slots_total + dst_slot) << A.n) + ampC;
#pragma unroll
    for (int j = 0; j < 16; ++j) LP_ST(dp + XSEL4R(j, g0, g1, g2, g3), a[j]);
  }
#undef LP_CA
#undef LP_CB
#undef LP_CC
}

// =====================================================================================
// layer sequence: k_layer_pass generalised to any order of the three aligned nibble sweeps and to
// XY pair rotations on two register bits (SeqPlan, pqc_common.cuh) -- the passes of the XXZ
// template, whose XY bonds alternate between nibbles.  Same arithmetic and op order as
// k_sweep_pass; the first sweep (tile positions 8-11 in registers) loads global memory
// directly; the last one stores directly when it holds positions 8-11 or 4-7, and through one
// more shared-memory transposition when it holds positions 0-3 (lanes must cover the low bits).
// =====================================================================================
struct SeqArgs {
  const c128* src;
  c128* dst;
  const double2* gtrig;
  int tstride, toff, ntrig;
  int nwt;
  int n;
  int lbit[12];
  int obit[PQC_MAX_QUBITS];
  int slots_total, active, nspawn;
  int spawn_slot[SEQ_MAX_SPAWN];
  double spawn_cr[SEQ_MAX_SPAWN], spawn_ci[SEQ_MAX_SPAWN];
  // staged ends (tile positions 0-2 are not the amplitude bits 0-2): global memory is accessed
  // in amplitude order -- lane bits = amplitude bits 0-3 -- and the tile is scattered into /
  // gathered from its swizzled shared-memory slots.  Thread bit t < 4 sits at tile position
  // st_q[t], thread bit 4 + t at st_hi[t]; register r contributes the masks st_rs / st_ra.
  int staged;
  int st_q[4], st_hi[4];
  uint32_t st_rs[4], st_ra[4];
  SeqPlan plan;
};


// Consecutive diagonal ops of a sweep (R_z phases, CZ signs) are accumulated per thread --
// thread-level phase T, one phase per register bit (bit = 1 takes the conjugate), a 16-bit sign
// mask -- and applied together: at most 16 complex multiplications per touched register bit
// whatever the number of gates.  `base` = the thread's tile-local index (register bits zero).
template <bool GEN, bool DIAG, bool LAYERS, int RN, int NX, int NY>
__device__ __forceinline__ void seq_ops(c128 (&a)[16], const SeqArgs& A, int slot, int vx, int vy,
                                        const double2* trig, const uint32_t (*s_wn)[3][16],
                                        const uint32_t* s_wb, int gen, double& fscale,
                                        uint32_t base, uint32_t tbase) {
  const int nops = A.plan.nops[slot], off = A.plan.off[slot];
  for (int oi = 0; oi < nops; ++oi) {
    const int kind = A.plan.ops[off + oi].kind;
    if (DIAG && (kind == PQC_OP_RZ || kind == PQC_OP_CZ)) {
      // ---- a run of diagonal ops: accumulate, then apply once
      double tc = 1.0, ts = 0.0;                  // thread-level phase (tc + i ts)
      double pc[4] = {1.0, 1.0, 1.0, 1.0}, pn[4] = {0.0, 0.0, 0.0, 0.0};   // bit k = 1: (pc + i pn)
      uint32_t sg = 0u, touched = 0u;             // sign mask over j; bit 4 of touched: T set
      for (; oi < nops; ++oi) {
        const FastOp& op = A.plan.ops[off + oi];
        if (op.kind == PQC_OP_RZ) {
          const double2 cs = trig[op.t[0]];
          const int k0 = op.t[1];
          if (k0 >= 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k == k0) {                      // (pc + i pn) *= (c + i s)
                const double c = pc[k], sn = pn[k];
                pc[k] = c * cs.x - sn * cs.y;
                pn[k] = sn * cs.x + c * cs.y;
              }
            touched |= 1u << k0;
          } else {
            const uint32_t bit = op.t[2] >= 0 ? ((base >> op.t[2]) & 1u) : ((tbase >> op.t[3]) & 1u);
            const double sn = bit ? cs.y : -cs.y, c = tc, s0 = ts;
            tc = c * cs.x - s0 * sn;
            ts = s0 * cs.x + c * sn;
            touched |= 16u;
          }
        } else if (op.kind == PQC_OP_CZ) {
          // 16-bit masks over j of "bit set": a register bit's pattern, else all / none
          auto mask_of = [&](int k, int l, int b) -> uint32_t {
            if (k >= 0) return k == 0 ? 0xAAAAu : (k == 1 ? 0xCCCCu : (k == 2 ? 0xF0F0u : 0xFF00u));
            const uint32_t bit = l >= 0 ? ((base >> l) & 1u) : ((tbase >> b) & 1u);
            return bit ? 0xFFFFu : 0u;
          };
          sg ^= mask_of(op.t[0], op.t[2], op.wt) & mask_of(op.t[1], op.t[3], op.nterms);
        } else {
          break;
        }
      }
      --oi;                                       // the outer loop steps past the run's last op
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (touched & (1u << k)) {                // warp-uniform
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const double sn = (j & (1 << k)) ? pn[k] : -pn[k];
            const c128 v = a[j];
            a[j] = make_double2(v.x * pc[k] - v.y * sn, v.y * pc[k] + v.x * sn);
          }
        }
      }
      if (touched & 16u) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const c128 v = a[j];
          a[j] = make_double2(v.x * tc - v.y * ts, v.y * tc + v.x * ts);
        }
      }
      if (sg) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if ((sg >> j) & 1u) a[j] = make_double2(-a[j].x, -a[j].y);
      }
      continue;
    }
    const FastOp& op = A.plan.ops[off + oi];
    if (kind == PQC_K_RXY) {
      const double2 cs = trig[op.t[0]];
      switch (op.subk) {
        case 1: op_xy<0, 1>(a, cs.x, cs.y); break;
        case 2: op_xy<0, 2>(a, cs.x, cs.y); break;
        case 3: op_xy<0, 3>(a, cs.x, cs.y); break;
        case 6: op_xy<1, 2>(a, cs.x, cs.y); break;
        case 7: op_xy<1, 3>(a, cs.x, cs.y); break;
        default: op_xy<2, 3>(a, cs.x, cs.y); break;
      }
    } else if (LAYERS && kind == PQC_K_LAYER_RX4) {
      const int sk = op.subk;
      double2 c0 = make_double2(0.0, 1.0), c1 = c0, c2 = c0, c3 = c0;
      if (sk & 0xff) c0 = trig[op.t[0]];
      if (sk & 0xff00) c1 = trig[op.t[1]];
      if (sk & 0xff0000) c2 = trig[op.t[2]];
      if (sk & 0xff000000) c3 = trig[op.t[3]];
      op_rx_t<0>(a, c0.x);
      op_rx_t<1>(a, c1.x);
      op_rx_t<2>(a, c2.x);
      op_rx_t<3>(a, c3.x);
      fscale *= (c0.y * c1.y) * (c2.y * c3.y);
    } else if (LAYERS && kind == PQC_K_LAYER_REAL4) {
      const int sk = op.subk;
      double f = 1.0;
#define SEQ_REAL_SLOT(K)                                                   \
  {                                                                        \
    const int kd = ((sk >> (8 * K)) & 0xff) - 1;                           \
    if (kd == PQC_OP_RY) { const double2 tc = trig[op.t[K]]; op_ry_t<K>(a, tc.x); f *= tc.y; } \
    else if (kd == PQC_OP_H) { op_h_u<K>(a); f *= 0.70710678118654752440; } \
  }
      SEQ_REAL_SLOT(0) SEQ_REAL_SLOT(1) SEQ_REAL_SLOT(2) SEQ_REAL_SLOT(3)
#undef SEQ_REAL_SLOT
      fscale *= f;
    } else {
      // ZZSUM / GEN: w(x) = w(tile) ^ w(thread nibbles) ^ w(register nibble value j)
      const uint32_t(*wn)[16] = s_wn[op.wt];
      const uint32_t w0 = s_wb[op.wt] ^ wn[NX][vx] ^ wn[NY][vy];
      const uint32_t w1 = wn[RN][1], w2 = wn[RN][2], w3 = wn[RN][4], w4 = wn[RN][8];
      if (kind == PQC_K_ZZSUM) {
        const double2* tz = trig + op.t[0];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const double2 ph = tz[__popc(w0 ^ XSEL4R(j, w1, w2, w3, w4))];
          const c128 v = a[j];
          a[j] = make_double2(v.x * ph.x - v.y * ph.y, v.y * ph.x + v.x * ph.y);
        }
      } else if (GEN && gen == op.spawn) {
        const double cr = A.spawn_cr[gen], ci = A.spawn_ci[gen];
        const int nt = op.nterms;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const double f = (double)(nt - 2 * __popc(w0 ^ XSEL4R(j, w1, w2, w3, w4)));
          const double fr = f * cr, fi = f * ci;
          const c128 v = a[j];
          a[j] = make_double2(v.x * fr - v.y * fi, v.y * fr + v.x * fi);
        }
      }
    }
  }
}

// LAYERS = false: the pass has no 4-slot rotation layer ops (every pass of the XXZ template): their
// branches are compiled out of the op loop (the same reasoning as k_tile_pipe's op sets)
template <bool GEN, bool DIAG, bool LAYERS>
__global__ void __launch_bounds__(256, 2) k_layer_seq(const SeqArgs A) {
  extern __shared__ __align__(16) c128 lp_sm[];
  double2* trig = reinterpret_cast<double2*>(lp_sm + 4096);
  __shared__ uint32_t s_ta[2][16];               // amplitude bits of tile nibbles 1 and 2
  __shared__ uint32_t s_wn[FAST_MAX_WT][3][16];  // linear forms per table, nibble, nibble value
  __shared__ uint32_t s_wb[FAST_MAX_WT];         // linear form of the tile index
  const int tid = threadIdx.x, lo = tid & 15, hi = tid >> 4;
  const int tiles_log2 = A.n - 12;
  const long long item = (long long)blockIdx.x >> tiles_log2;
  const uint32_t tile = (uint32_t)(blockIdx.x & ((1u << tiles_log2) - 1u));
  uint32_t tbase = 0;
  for (int j = 0; j < tiles_log2; ++j) tbase |= ((tile >> j) & 1u) << A.obit[j];
  const int ips = A.active + A.nspawn;
  const long long sample = item / ips;
  const int r = (int)(item - sample * ips);
  int src_slot = r, dst_slot = r, gen = -1;
  if (GEN && r >= A.active) {
    src_slot = 0;
    gen = r - A.active;
    dst_slot = A.spawn_slot[gen];
  }
  // amplitude bits fed by `lo` (tile positions 0-3; position 3 need not be amplitude bit 3)
  const uint32_t lo_amp = ((uint32_t)lo & 7u) | ((((uint32_t)lo >> 3) & 1u) << A.lbit[3]);
  // staged ends: this thread's tile index / amplitude offset in amplitude order
  uint32_t st_sm = 0, st_amp = 0;
  if (A.staged) {
    uint32_t bi = 0;
    st_amp = tbase | (uint32_t)lo;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      bi |= (((uint32_t)lo >> t) & 1u) << A.st_q[t];
      bi |= (((uint32_t)hi >> t) & 1u) << A.st_hi[t];
      st_amp |= (((uint32_t)hi >> t) & 1u) << A.lbit[A.st_hi[t]];
    }
    st_sm = swz(bi);
  }
  // ---- sweep 0 loads straight from global memory: registers = tile positions 8-11
  c128 a[16];
  if (A.staged) {
    const c128* sp = A.src + ((sample * A.slots_total + src_slot) << A.n) + st_amp;
    const uint32_t r0 = A.st_ra[0], r1 = A.st_ra[1], r2 = A.st_ra[2], r3 = A.st_ra[3];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = sp[XSEL4R(j, r0, r1, r2, r3)];
  } else {
    const uint32_t ampA = tbase | lo_amp | ((uint32_t)(hi & 1) << A.lbit[4]) |
                          ((uint32_t)((hi >> 1) & 1) << A.lbit[5]) |
                          ((uint32_t)((hi >> 2) & 1) << A.lbit[6]) |
                          ((uint32_t)((hi >> 3) & 1) << A.lbit[7]);
    const uint32_t g0 = 1u << A.lbit[8], g1 = 1u << A.lbit[9], g2 = 1u << A.lbit[10],
                   g3 = 1u << A.lbit[11];
    const c128* sp = A.src + ((sample * A.slots_total + src_slot) << A.n) + ampA;
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = sp[XSEL4R(j, g0, g1, g2, g3)];
  }
  // ---- tables (overlap the tile load)
  if (tid < 32) {
    uint32_t v = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) v |= (((uint32_t)lo >> i) & 1u) << A.lbit[4 + 4 * hi + i];
    s_ta[hi][lo] = v;
  }
  for (int e = tid; e < A.nwt * 48; e += 256) (&s_wn[0][0][0])[e] = (&A.plan.wn[0][0][0])[e];
  if (tid < A.nwt) {
    uint32_t w = 0;
    for (int j = 0; j < tiles_log2; ++j) w ^= ((tile >> j) & 1u) ? A.plan.wo[tid][j] : 0u;
    s_wb[tid] = w;
  }
  for (int e = tid; e < A.ntrig; e += 256) trig[e] = A.gtrig[sample * A.tstride + A.toff + e];
  if (A.staged) {
    const uint32_t q0 = A.st_rs[0], q1 = A.st_rs[1], q2 = A.st_rs[2], q3 = A.st_rs[3];
#pragma unroll
    for (int j = 0; j < 16; ++j) lp_sm[st_sm ^ XSEL4R(j, q0, q1, q2, q3)] = a[j];
  }
  __syncthreads();

  double fscale = 1.0;
#define LP_CA(j) ((((j) << 8) ^ ((((j) << 2) ^ ((j) >> 1)) & 7)))   /* swz(j << 8) */
#define LP_CB(j) (((j) ^ ((j) >> 3)))                               /* swz(j)      */
#define LP_CC(j) ((((j) << 4) ^ ((((j) << 1) ^ ((j) >> 2)) & 7)))   /* swz(j << 4) */
  const uint32_t sbA = swz((uint32_t)tid), sbB = swz((uint32_t)tid << 4),
                 sbC = swz((uint32_t)lo | ((uint32_t)hi << 8));
  const int nsw = A.plan.nsw;
  int g = 0;
  for (int s = 0; s < nsw; ++s) {
    g = A.plan.geom[s];
    if (s > 0 || A.staged) {
      if (g == 0) {
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = lp_sm[sbA ^ LP_CA(j)];
      } else if (g == 1) {
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = lp_sm[sbB ^ LP_CB(j)];
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = lp_sm[sbC ^ LP_CC(j)];
      }
    }
    if (g == 0)
      seq_ops<GEN, DIAG, LAYERS, 2, 0, 1>(a, A, s, lo, hi, trig, s_wn, s_wb, gen, fscale, (uint32_t)tid, tbase);
    else if (g == 1)
      seq_ops<GEN, DIAG, LAYERS, 0, 1, 2>(a, A, s, lo, hi, trig, s_wn, s_wb, gen, fscale, (uint32_t)tid << 4, tbase);
    else
      seq_ops<GEN, DIAG, LAYERS, 1, 0, 2>(a, A, s, lo, hi, trig, s_wn, s_wb, gen, fscale,
                            (uint32_t)lo | ((uint32_t)hi << 8), tbase);
    if (s + 1 == nsw && g != 1 && !A.staged) break;   // the registers go straight to global memory
    if (s + 1 == nsw && fscale != 1.0) op_scale(a, fscale);
    // back to shared memory (every thread rewrites exactly the slots it read, except after the
    // global load of sweep 0, which nobody has touched yet)
    if (g == 0) {
#pragma unroll
      for (int j = 0; j < 16; ++j) lp_sm[sbA ^ LP_CA(j)] = a[j];
    } else if (g == 1) {
#pragma unroll
      for (int j = 0; j < 16; ++j) lp_sm[sbB ^ LP_CB(j)] = a[j];
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) lp_sm[sbC ^ LP_CC(j)] = a[j];
    }
    __syncthreads();
  }
  c128* dbase = A.dst + ((sample * A.slots_total + dst_slot) << A.n);
  if (A.staged) {
    // the finished tile is in shared memory: gather it in amplitude order
    const uint32_t q0 = A.st_rs[0], q1 = A.st_rs[1], q2 = A.st_rs[2], q3 = A.st_rs[3];
    const uint32_t r0 = A.st_ra[0], r1 = A.st_ra[1], r2 = A.st_ra[2], r3 = A.st_ra[3];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = lp_sm[st_sm ^ XSEL4R(j, q0, q1, q2, q3)];
#pragma unroll
    for (int j = 0; j < 16; ++j) dbase[st_amp | XSEL4R(j, r0, r1, r2, r3)] = a[j];
  } else if (g == 0) {
    if (fscale != 1.0) op_scale(a, fscale);
    const uint32_t ampA = tbase | lo_amp | s_ta[0][hi];
    const uint32_t g0 = 1u << A.lbit[8], g1 = 1u << A.lbit[9], g2 = 1u << A.lbit[10],
                   g3 = 1u << A.lbit[11];
#pragma unroll
    for (int j = 0; j < 16; ++j) dbase[ampA | XSEL4R(j, g0, g1, g2, g3)] = a[j];
  } else {
    if (g == 1) {
      // the last sweep held positions 0-3 (already scaled and written back): one more
      // transposition so that lanes cover the low index bits
#pragma unroll
      for (int j = 0; j < 16; ++j) a[j] = lp_sm[sbC ^ LP_CC(j)];
    } else if (fscale != 1.0) {
      op_scale(a, fscale);
    }
    const uint32_t ampC = tbase | lo_amp | s_ta[1][hi];
    const uint32_t g0 = 1u << A.lbit[4], g1 = 1u << A.lbit[5], g2 = 1u << A.lbit[6],
                   g3 = 1u << A.lbit[7];
#pragma unroll
    for (int j = 0; j < 16; ++j) dbase[ampC | XSEL4R(j, g0, g1, g2, g3)] = a[j];
  }
#undef LP_CA
#undef LP_CB
#undef LP_CC
}

// =====================================================================================
// tile gather: dst slot <- (sum of Pauli terms) src slot 0, for one or more parameters.
// One CTA per (sample, 4096-amplitude chunk); the chunk of psi sits in shared memory, terms
// flipping only low bits read it from there, the others from global / L2.
// =====================================================================================
struct GatherArgs {
  c128* buf;
  int n, cb, slots_total;
  int nparams;
  int slot[V1_MAX_SPAWN], goff[V1_MAX_SPAWN], gcnt[V1_MAX_SPAWN];
  unsigned char purex[V1_MAX_SPAWN];   // every term is an X-string with one common coefficient
  const GenTerm* gens;
};

__device__ __forceinline__ double flip_sign(double v, uint32_t s31) {
  return __hiloint2double(__double2hiint(v) ^ (int)s31, __double2loint(v));
}

__global__ void __launch_bounds__(256, 2) k_tile_gather(const GatherArgs A) {
  extern __shared__ __align__(16) unsigned char smraw[];
  c128* sm = reinterpret_cast<c128*>(smraw);
  __shared__ uint32_t s_xm[64];
  const int chunks_log2 = A.n - A.cb;
  const long long s = blockIdx.x >> chunks_log2;
  const uint32_t c0 = (uint32_t)(blockIdx.x & ((1ll << chunks_log2) - 1)) << A.cb;
  const uint32_t csize = 1u << A.cb;
  const c128* psi = A.buf + ((s * A.slots_total) << A.n);
  for (uint32_t i = threadIdx.x; i < csize; i += 256) sm[i] = psi[c0 + i];
  __syncthreads();
  for (int p = 0; p < A.nparams; ++p) {
    c128* out = A.buf + ((s * A.slots_total + A.slot[p]) << A.n);
    const GenTerm* terms = A.gens + A.goff[p];
    const int nt = A.gcnt[p];
    // 16 outputs per thread (element e = tid + 256 r); terms visited once each
    for (uint32_t i0 = threadIdx.x; i0 < csize; i0 += 256 * 16) {
      double are[16], aim[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) are[r] = aim[r] = 0.0;
      if (A.purex[p]) {
        // sum_t psi[x ^ xmask_t], scaled once at the end: 2 adds per term and amplitude.
        // The masks are staged in shared memory so no iteration waits on a global load.
        __syncthreads();
        for (int t = threadIdx.x; t < nt && t < 64; t += 256) s_xm[t] = terms[t].xmask;
        __syncthreads();
        for (int t = 0; t < nt; ++t) {
          const uint32_t xm = t < 64 ? s_xm[t] : terms[t].xmask;
          if (xm >> A.cb) {
            const uint32_t xb = (c0 + i0) ^ xm;
#pragma unroll
            for (int r = 0; r < 16; ++r)
              if (i0 + 256u * r < csize) {
                const c128 v = psi[xb ^ (256u * r)];
                are[r] += v.x;
                aim[r] += v.y;
              }
          } else {
            const uint32_t b = i0 ^ (xm & 255u), hx = xm >> 8;
#pragma unroll
            for (int r = 0; r < 16; ++r)
              if (i0 + 256u * r < csize) {
                const c128 v = sm[b + 256u * ((uint32_t)r ^ hx)];
                are[r] += v.x;
                aim[r] += v.y;
              }
          }
        }
        const double cr = terms[0].re, ci = terms[0].im;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const uint32_t i = i0 + 256u * r;
          if (i < csize) out[c0 + i] = make_double2(are[r] * cr - aim[r] * ci, are[r] * ci + aim[r] * cr);
        }
      } else {
        for (int t = 0; t < nt; ++t) {
          const GenTerm g = terms[t];
          const bool far = (g.xmask >> A.cb) != 0;
          // coefficient times i^(number of Y factors)
          const int ny = __popc(g.xmask & g.zmask) & 3;
          double cr = g.re, ci = g.im;
          if (ny == 1) { cr = -g.im; ci = g.re; }
          else if (ny == 2) { cr = -g.re; ci = -g.im; }
          else if (ny == 3) { cr = g.im; ci = -g.re; }
#pragma unroll
          for (int r = 0; r < 16; ++r) {
            const uint32_t i = i0 + 256u * r;
            if (i >= csize) continue;
            const uint32_t x = (c0 + i) ^ g.xmask;
            const c128 v = far ? psi[x] : sm[x - c0];
            const uint32_t sg = ((uint32_t)__popc(x & g.zmask) & 1u) << 31;
            const double vx = flip_sign(v.x, sg), vy = flip_sign(v.y, sg);
            are[r] = fma(cr, vx, fma(-ci, vy, are[r]));
            aim[r] = fma(cr, vy, fma(ci, vx, aim[r]));
          }
        }
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const uint32_t i = i0 + 256u * r;
          if (i < csize) out[c0 + i] = make_double2(are[r], aim[r]);
        }
      }
    }
  }
}

// Fast path of the gather for X-string generators with one common coefficient (the R_x / R_xx
// layers of the Hamiltonian-variational templates), n >= 12: out = c * sum_t psi[x ^ xmask_t].
// One CTA per (sample, 4096-amplitude chunk); thread tid owns elements tid + 256 r, r < 16,
// r < 16.  Single-bit masks inside the chunk read it from shared memory with immediate offsets
// (one LDS + two adds per amplitude), masks that leave the chunk read psi through L2.
#define XS_TB 8                   // thread bits: 256 threads x 16 outputs (512 x 8 measured 12 % slower)
#define XS_NT (1 << XS_TB)
#define XS_R (4096 / XS_NT)
__global__ void __launch_bounds__(XS_NT, 2) k_xsum_gather(const GatherArgs A) {
  extern __shared__ __align__(16) c128 xs_sm[];
  __shared__ uint32_t s_xm[64];
  const int tid = threadIdx.x;
  const int chunks_log2 = A.n - 12;
  const long long s = blockIdx.x >> chunks_log2;
  const uint32_t c0 = (uint32_t)(blockIdx.x & ((1ll << chunks_log2) - 1)) << 12;
  const c128* psi = A.buf + ((s * A.slots_total) << A.n);
  {
    c128 own[XS_R];
#pragma unroll
    for (int r = 0; r < XS_R; ++r) own[r] = psi[c0 + tid + XS_NT * r];
#pragma unroll
    for (int r = 0; r < XS_R; ++r) xs_sm[tid + XS_NT * r] = own[r];
  }
  for (int p = 0; p < A.nparams; ++p) {
    const GenTerm* terms = A.gens + A.goff[p];
    const int nt = A.gcnt[p];
    __syncthreads();
    for (int t = tid; t < nt && t < 64; t += XS_NT) s_xm[t] = terms[t].xmask;
    __syncthreads();
    double are[XS_R], aim[XS_R];
#pragma unroll
    for (int r = 0; r < XS_R; ++r) are[r] = aim[r] = 0.0;
    for (int t = 0; t < nt; ++t) {
      const uint32_t xm = t < 64 ? s_xm[t] : terms[t].xmask;
      const uint32_t lo = xm & (XS_NT - 1u), hx = (xm >> XS_TB) & (XS_R - 1u), far = xm >> 12;
      if (far == 0 && lo == 0 && (hx & (hx - 1)) == 0) {
        // one of the thread's own register bits: immediate offsets
#define XS_OWN(K)                                                              \
  _Pragma("unroll") for (int r = 0; r < XS_R; ++r) {                           \
    const c128 v = xs_sm[tid + XS_NT * (r ^ (1 << (K)))];                      \
    are[r] += v.x;                                                             \
    aim[r] += v.y;                                                             \
  }
        if (hx == 1) { XS_OWN(0) } else if (hx == 2) { XS_OWN(1) } else if (hx == 4) { XS_OWN(2) }
        else { XS_OWN((XS_R == 16 ? 3 : 2)) }
#undef XS_OWN
      } else if (far == 0 && hx == 0) {
        const c128* b = xs_sm + (tid ^ lo);
#pragma unroll
        for (int r = 0; r < XS_R; ++r) {
          const c128 v = b[XS_NT * r];
          are[r] += v.x;
          aim[r] += v.y;
        }
      } else if (far == 0) {
        const uint32_t b = (uint32_t)tid ^ lo;
#pragma unroll
        for (int r = 0; r < XS_R; ++r) {
          const c128 v = xs_sm[b + XS_NT * ((uint32_t)r ^ hx)];
          are[r] += v.x;
          aim[r] += v.y;
        }
      } else if ((xm & 0xfffu) == 0) {
        const c128* b = psi + ((c0 ^ xm) + tid);
#pragma unroll
        for (int r = 0; r < XS_R; ++r) {
          const c128 v = b[XS_NT * r];
          are[r] += v.x;
          aim[r] += v.y;
        }
      } else {
        const uint32_t xb = (c0 + tid) ^ xm;
#pragma unroll
        for (int r = 0; r < XS_R; ++r) {
          const c128 v = psi[xb ^ ((uint32_t)XS_NT * r)];
          are[r] += v.x;
          aim[r] += v.y;
        }
      }
    }
    const double cr = terms[0].re, ci = terms[0].im;
    c128* out = A.buf + ((s * A.slots_total + A.slot[p]) << A.n) + c0 + tid;
#pragma unroll
    for (int r = 0; r < XS_R; ++r)
      out[XS_NT * r] = make_double2(are[r] * cr - aim[r] * ci, are[r] * ci + aim[r] * cr);
  }
}

// A thread-block-cluster form of this kernel (the partner chunks of up to four far bits read from
// the peer CTAs' shared memory instead of L2) was measured and dropped: bit-identical results, but
// the headline step lost 1.8 % at cluster size 2 and 14 % at 16 -- DSMEM reads are slower than the L2
// hits they replace (profiles/r2_xsum_cluster.md).

// standalone Gram columns: one CTA per (sample, row j, partner); writes the sum into tile 0 of
// the partial array and zeroes the other tiles so the reducer stays uniform.
__global__ void __launch_bounds__(256) k_multi_dots(const c128* __restrict__ buf, int n,
                                                    int slots_total, int rows, int npart,
                                                    V1Args A) {
  __shared__ double red[32];
  const long long D = 1ll << n;
  long long b = blockIdx.x;
  const int q = (int)(b % npart);
  b /= npart;
  const int j = (int)(b % rows);
  const long long s = b / rows;
  const int ps = A.partner_slot[q];
  if (j > ps) return;
  const c128* x = buf + ((s * slots_total + j) << n);
  const c128* y = buf + ((s * slots_total + ps) << n);
  double re = 0.0, im = 0.0;
  for (long long i = threadIdx.x; i < D; i += 256) {
    const c128 u = x[i], v = y[i];
    re += u.x * v.x + u.y * v.y;
    im += u.x * v.y - u.y * v.x;
  }
  re = block_sum<256>(re, red);
  im = block_sum<256>(im, red);
  if (threadIdx.x == 0) {
    c128* g = A.gpart + ((s * (A.P + 1) + j) * A.P + (ps - 1)) * A.ntiles;
    g[0] = make_double2(re, im);
    for (int t = 1; t < A.ntiles; ++t) g[t] = make_double2(0.0, 0.0);
  }
}

// F_pq = 4 Re(G_pq - conj(s_p) s_q), p <= q, mirrored (measure.py:55-70), G summed over tiles
// in a fixed order (bitwise reproducible).
__global__ void k_qfim_reduce(const c128* __restrict__ gpart, long long S, int P, int ntiles,
                              double* __restrict__ F) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= S * P * P) return;
  const long long s = e / ((long long)P * P);
  const int r = (int)((e / P) % P), c = (int)(e % P);
  const int p = r < c ? r : c, q = r < c ? c : r;
  auto G = [&](int row, int col) -> c128 {
    const c128* g = gpart + ((s * (P + 1) + row) * P + col) * ntiles;
    double re = 0.0, im = 0.0;
    for (int t = 0; t < ntiles; ++t) { re += g[t].x; im += g[t].y; }
    return make_double2(re, im);
  };
  const c128 sp = G(0, p), sq = G(0, q), d = G(1 + p, q);
  F[e] = 4.0 * (d.x - (sp.x * sq.x + sp.y * sq.y));
}

// =====================================================================================
// QFIM Gram matrix on the FP64 tensor cores.  measure.py:55-70 needs only
//   F_pq = 4 (Re<d_p|d_q> - Re(conj<psi|d_p> <psi|d_q>)),
// and Re<d_p|d_q> is the REAL inner product of the two vectors read as 2 D doubles.  So the
// P x P block is one real symmetric Gram V^T V over K = 2 D (half the DMMAs of the complex
// product, and no padding row for psi), while the P complex overlaps <psi|d_p> are taken by
// the CUDA cores from the same shared-memory stages.  All vectors are taken at ONE common
// time (overlaps are invariant under the later unitary gates).
// One CTA per (parameter set, K slice of 1024 amplitudes); warp (kw, grp) owns the k4-steps
// kw, kw + ksh, .. of every 64-amplitude stage for the upper-triangular 8x8 tiles
// [grp*16, grp*16+16).  Stages arrive by TMA: one cp.async.bulk of 1 KB per vector row,
// completion counted on an mbarrier per ring slot, so no thread spends issue slots on the
// copies; the K shares and the K slices are added in a fixed order (bitwise reproducible).
// Rows 0..PF-1 are slots 1.. of `buf` (slot 0 = psi), rows PF.. are slots 1.. of `buf2`.
// =====================================================================================
#define GR_K 64                   // amplitudes per stage: 1 KB contiguous per vector row
#define GR_NS_MAX 4               // TMA ring depth (ns - 1 stages in flight while one is used)
#define GR_ROW (GR_K + 2)         // padded complex per smem row: the 8 rows of a fragment load
                                  // (8 B per lane) fall into distinct banks
#define GR_TPW 16                 // 8x8 output tiles per warp (one "tile group")
#define GR_SLICE 1024             // amplitudes per CTA
#define GR_MAXSPLIT 64
#define GR_PSIROWS 9              // <psi|d_p> rows per warp (P <= 72, >= 8 warps)

__device__ __forceinline__ void gr_dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

static int gram_ksplit(int n) {
  const long long D = 1ll << n;
  return (int)std::max<long long>(1, std::min<long long>(GR_MAXSPLIT, D / GR_SLICE));
}

// SMALL: P <= 32 (at most 4 row blocks, 10 tiles, one tile group of 8 K-share warps): the 4
// row-block fragments of a k4-step are loaded once and feed all tiles from registers.

// SMALL: P <= 32 (at most 4 row blocks, 10 tiles, one tile group of 8 K-share warps): the 4
// row-block fragments of a k4-step are loaded once and feed all tiles from registers.
template <int NTMAX, int MINB, bool SMALL>
__global__ void __launch_bounds__(NTMAX + 32, MINB) k_gram_real(const c128* __restrict__ buf, int n,
                                                           int slots_total, int PF,
                                                           const c128* __restrict__ buf2, int slots2,
                                                           int P, int P8, int ksplit, int ksh,
                                                           int ns, double* __restrict__ gpart) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ __align__(8) uint64_t full[GR_NS_MAX], empty[GR_NS_MAX];
  // the last warp is the TMA producer; the others are consumers
  const int NT = blockDim.x - 32, nw = NT >> 5;
  c128* sm = reinterpret_cast<c128*>(smraw);                       // [ns][P8 + 1][GR_ROW]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const bool producer = warp == nw;
  const int kw = warp % ksh, grp = warp / ksh;
  const long long s = blockIdx.x / ksplit;
  const int ks = blockIdx.x % ksplit;
  const long long D = 1ll << n, kbeg = (D / ksplit) * ks;
  const int nk = (int)((D / ksplit) / GR_K);
  const c128* V = buf + ((s * slots_total) << n);
  const c128* V2 = buf2 + (s * slots2 + 1 - PF) * D;   // row r >= PF: V2 + r D
  const int rows = P8 + 1;                     // smem rows per stage; row P8 holds psi
  const int T = P8 / 8, ntile = T * (T + 1) / 2;
  const int t_begin = grp * GR_TPW, t_end = min(ntile, t_begin + GR_TPW);
  // tile index -> (row block, col block), 4 bits each, upper triangle row-major
  unsigned long long tij_lo = 0, tij_hi = 0;   // 16 tiles x 8 bits
  if (!SMALL) {
    int idx = 0;
    for (int i = 0; i < T; ++i)
      for (int j = i; j < T; ++j, ++idx)
        if (idx >= t_begin && idx < t_end) {
          const int q = idx - t_begin;
          const unsigned long long v = (unsigned long long)(i | (j << 4)) << (8 * (q & 7));
          if (q < 8) tij_lo |= v; else tij_hi |= v;
        }
  }
  // pad rows P..P8-1 of every ring slot stay zero (TMA only writes rows < P and the psi row)
  for (int e = tid; e < ns * (P8 - P) * GR_ROW; e += blockDim.x) {
    const int b = e / ((P8 - P) * GR_ROW), rem = e - b * (P8 - P) * GR_ROW;
    sm[(b * rows + P) * GR_ROW + rem] = make_double2(0.0, 0.0);
  }
  if (tid == 0) {
    for (int i = 0; i < ns; ++i) {
      gr_mbar_init(&full[i], 1);
      gr_mbar_init(&empty[i], nw);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  double acc[GR_TPW][2];
#pragma unroll
  for (int q = 0; q < GR_TPW; ++q) acc[q][0] = acc[q][1] = 0.0;
  double sre[GR_PSIROWS], sim[GR_PSIROWS];
#pragma unroll
  for (int m = 0; m < GR_PSIROWS; ++m) sre[m] = sim[m] = 0.0;
  if (producer) {
    // runs ahead of the consumers by up to ns stages; lane l copies the rows l, l + 32, ..
    for (int it = 0; it < nk; ++it) {
      const int b = it % ns;
      if (it >= ns) gr_mbar_wait(&empty[b], (unsigned)(((it / ns) - 1) & 1));
      const long long k0 = kbeg + (long long)it * GR_K;
      if (lane == 0) gr_mbar_expect(&full[b], (unsigned)((P + 1) * GR_K * sizeof(c128)));
      __syncwarp();
      for (int r = lane; r <= P; r += 32) {
        const c128* src = r == P ? V : (r < PF ? V + ((long long)(r + 1) << n)
                                               : V2 + ((long long)r << n));
        gr_bulk_load(sm + (b * rows + (r == P ? P8 : r)) * GR_ROW, src + k0,
                     (unsigned)(GR_K * sizeof(c128)), &full[b]);
      }
    }
  } else {
    for (int it = 0; it < nk; ++it) {
      const int b = it % ns;
      gr_mbar_wait(&full[b], (unsigned)((it / ns) & 1));
      // real Gram: k4-step kk covers the 4 doubles (2 amplitudes) 4 kk .. 4 kk + 3 of every row
      for (int kk = kw; kk < GR_K / 2; kk += ksh) {
        const double* a_s = reinterpret_cast<const double*>(sm + (b * rows + g) * GR_ROW) + 4 * kk + t4;
        if (SMALL) {
          double f[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) f[i] = i < T ? a_s[i * (8 * GR_ROW * 2)] : 0.0;
          int q = 0;
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = i; j < 4; ++j, ++q)
              if (j < T) gr_dmma(acc[q][0], acc[q][1], f[i], f[j]);     // warp-uniform
        } else {
#pragma unroll
          for (int q = 0; q < GR_TPW; ++q) {
            if (t_begin + q < t_end) {          // warp-uniform
              const unsigned ij = (unsigned)(((q < 8 ? tij_lo : tij_hi) >> (8 * (q & 7))) & 0xff);
              const double fa = a_s[(ij & 15u) * (8 * GR_ROW * 2)];
              const double fb = a_s[(ij >> 4) * (8 * GR_ROW * 2)];
              gr_dmma(acc[q][0], acc[q][1], fa, fb);
            }
          }
        }
      }
      // <psi|d_p> for the rows p = warp, warp + nw, ..: two amplitudes of the stage per lane
      {
        const c128* ps = sm + (b * rows + P8) * GR_ROW;
        const c128 y0 = ps[lane], y1 = ps[lane + 32];
#pragma unroll
        for (int m = 0; m < GR_PSIROWS; ++m) {
          const int p = warp + nw * m;
          if ((!SMALL || m < 4) && p < P) {
            const c128* d = sm + (b * rows + p) * GR_ROW;
            const c128 x0 = d[lane], x1 = d[lane + 32];
            sre[m] += y0.x * x0.x + y0.y * x0.y + y1.x * x1.x + y1.y * x1.y;
            sim[m] += y0.x * x0.y - y0.y * x0.x + y1.x * x1.y - y1.y * x1.x;
          }
        }
      }
      __syncwarp();
      if (lane == 0) gr_mbar_arrive(&empty[b]);   // this warp is done with the slot
    }
  }
  __syncthreads();
  // add the K shares of every tile group in a fixed order through the (now idle) stage
  // buffers: scratch [grp][q][2][32] doubles
  double* scr = reinterpret_cast<double*>(sm);
  for (int r = 1; r < ksh; ++r) {
    if (!producer && kw == r) {
      double* o = scr + (size_t)grp * GR_TPW * 2 * 32 + lane;
#pragma unroll
      for (int q = 0; q < GR_TPW; ++q) {
        o[(q * 2 + 0) * 32] = acc[q][0];
        o[(q * 2 + 1) * 32] = acc[q][1];
      }
    }
    __syncthreads();
    if (!producer && kw == 0) {
      const double* o = scr + (size_t)grp * GR_TPW * 2 * 32 + lane;
#pragma unroll
      for (int q = 0; q < GR_TPW; ++q) {
        acc[q][0] += o[(q * 2 + 0) * 32];
        acc[q][1] += o[(q * 2 + 1) * 32];
      }
    }
    __syncthreads();
  }
  double* out = gpart + (s * ksplit + ks) * ((long long)P * P + 2 * P);
  if (!producer && kw == 0) {
    if (SMALL) {
      int q = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i; j < 4; ++j, ++q)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int row = 8 * i + g, col = 8 * j + 2 * t4 + c;
            if (row < P && col < P) out[(long long)row * P + col] = acc[q][c];
          }
    } else {
#pragma unroll
      for (int q = 0; q < GR_TPW; ++q)
        if (t_begin + q < t_end) {
          const unsigned ij = (unsigned)(((q < 8 ? tij_lo : tij_hi) >> (8 * (q & 7))) & 0xff);
          const int ti = ij & 15, tj = ij >> 4;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int row = 8 * ti + g, col = 8 * tj + 2 * t4 + c;
            if (row < P && col < P) out[(long long)row * P + col] = acc[q][c];
          }
        }
    }
  }
#pragma unroll
  for (int m = 0; m < GR_PSIROWS; ++m) {
    const int p = warp + nw * m;
    if (!producer && (!SMALL || m < 4) && p < P) {          // warp-uniform
      const double re = warp_sum(sre[m]), im = warp_sum(sim[m]);
      if (lane == 0) {
        out[(long long)P * P + p] = re;
        out[(long long)P * P + P + p] = im;
      }
    }
  }
}

// F_pq = 4 (R[p][q] - Re(conj(s_p) s_q)), p <= q, mirrored; K slices summed in a fixed order.
// `inv` (meet-in-the-middle plan): parameter -> Gram row, sign bit set for the vectors of the
// backward pipeline, which carry -d_p (their gates run with negated angles).
__global__ void k_qfim_from_gram(const double* __restrict__ gpart, long long S, int P, int ksplit,
                                 const int* __restrict__ inv, double* __restrict__ F) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= S * P * P) return;
  const long long s = e / ((long long)P * P);
  int r = (int)((e / P) % P), c = (int)(e % P);
  double sign = 1.0;
  if (inv) {
    const int ir = inv[r], ic = inv[c];
    if ((ir ^ ic) < 0) sign = -1.0;
    r = ir & 0x7fffffff;
    c = ic & 0x7fffffff;
  }
  const int p = r < c ? r : c, q = r < c ? c : r;
  const long long stride = (long long)P * P + 2 * P;
  double R = 0.0, pr = 0.0, pi = 0.0, qr = 0.0, qi = 0.0;
  for (int ks = 0; ks < ksplit; ++ks) {
    const double* o = gpart + (s * ksplit + ks) * stride;
    R += o[(long long)p * P + q];
    pr += o[(long long)P * P + p];
    pi += o[(long long)P * P + P + p];
    qr += o[(long long)P * P + q];
    qi += o[(long long)P * P + P + q];
  }
  F[e] = sign * 4.0 * (R - (pr * qr + pi * qi));
}

bool pqc_v1_gram_ok(const pqc_program* prog) {
  static int off = -1;
  if (off < 0) {
    const char* e = getenv("PQC_QFIM_GRAM");
    off = (e && strcmp(e, "0") == 0) ? 1 : 0;
  }
  // T <= 9 row blocks (P <= 72): 45 tiles = 3 tile groups of 16
  return !off && prog->n >= 7 && prog->P >= 1 && (prog->P + 7) / 8 <= 9;
}

int pqc_v1_gram_qfim(const pqc_program* prog, const c128* buf, long long S, c128* d_gpart,
                     double* d_F, cudaStream_t st) {
  return pqc_v1_gram_qfim2(prog, buf, prog->P + 1, prog->P + 1, buf, prog->P + 1, nullptr, S,
                           d_gpart, d_F, st);
}

// M1 = rows taken from `buf` including psi (slot 0); the other derivative rows are slots 1..
// of `buf2`.
int pqc_v1_gram_qfim2(const pqc_program* prog, const c128* buf, int slots1, int M1,
                      const c128* buf2, int slots2, const int* d_inv, long long S, c128* d_gpart,
                      double* d_F, cudaStream_t st) {
  const int P = prog->P, P8 = (P + 7) & ~7, PF = M1 - 1;
  const int T = P8 / 8, ntile = T * (T + 1) / 2, ngrp = (ntile + GR_TPW - 1) / GR_TPW;
  const int ksh = ngrp == 1 ? 8 : 4;            // K shares: at least 8 warps per CTA
  const int nthreads = 32 * ksh * ngrp;
  const int ksplit = gram_ksplit(prog->n);
  const size_t stage = (size_t)(P8 + 1) * GR_ROW * sizeof(c128);
  // ring depth: P <= 32 runs 2 CTAs per SM with 3 slots each, larger P one CTA with 2-4 slots
  const int ns = T <= 4 ? 3 : (int)std::max<size_t>(2, std::min<size_t>(GR_NS_MAX, (200 * 1024) / stage));
  const size_t smem = std::max((size_t)ns * stage, (size_t)ngrp * GR_TPW * 2 * 32 * sizeof(double));
  static PqcDeviceOnce attr_once;
  if (attr_once.first()) {
    PQC_CUDA(cudaFuncSetAttribute(k_gram_real<256, 2, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    PQC_CUDA(cudaFuncSetAttribute(k_gram_real<384, 1, false>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  if (S * ksplit > 0x7fffffffLL) PQC_FAIL(-1, "gram grid too large");
  double* gp = reinterpret_cast<double*>(d_gpart);
  if (T <= 4)
    k_gram_real<256, 2, true><<<(unsigned)(S * ksplit), nthreads + 32, smem, st>>>(
        buf, prog->n, slots1, PF, buf2, slots2, P, P8, ksplit, ksh, ns, gp);
  else
    k_gram_real<384, 1, false><<<(unsigned)(S * ksplit), nthreads + 32, smem, st>>>(
        buf, prog->n, slots1, PF, buf2, slots2, P, P8, ksplit, ksh, ns, gp);
  PQC_LAUNCH_CHECK();
  const long long tot = S * P * P;
  k_qfim_from_gram<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(gp, S, P, ksplit, d_inv, d_F);
  PQC_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================
// host drivers
// =====================================================================================
static int fill_pass_args(const pqc_program* prog, const V1Pass& ps, V1Args& a) {
  a.mops = prog->d_mops + ps.mop_off;
  a.sweeps = prog->d_sweeps + ps.sweep_off;
  a.nsweeps = ps.nsweeps;
  a.sweeps_nmops = ps.nmops;
  a.sweep0_io = ps.io_first;
  a.last_io = ps.io_last;
  a.gtrig = prog->d_trig;
  a.toff = ps.trig_goff;
  a.ntrig = ps.ntrig;
  a.wtab = prog->d_zz + ps.wt_off;
  a.nwt = ps.nwt;
  a.hpass = &ps;
  a.hprog = prog;
  a.gens = prog->d_gens;
  a.n = prog->n;
  a.tb = ps.tb;
  a.items_log2 = V1_LOCAL_BITS - ps.tb;
  a.low_run = ps.low_run;
  memcpy(a.lbit, ps.lbit, sizeof(a.lbit));
  memcpy(a.obit, ps.obit, sizeof(a.obit));
  return 0;
}

static int launch_init(c128* buf, int mode, const c128* init, long long init_stride, long long S,
                       int slots_total, int n, cudaStream_t st) {
  const long long total = S << n;
  const long long grid = std::min<long long>((total + 255) / 256, 148 * 16);
  k_init_slot0<<<(unsigned)grid, 256, 0, st>>>(buf, mode, init, init_stride, S, slots_total, n);
  PQC_LAUNCH_CHECK();
  return 0;
}

static int prefetch_dist() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PQC_PREFETCH");
    v = e ? atoi(e) : 0;
  }
  return v;
}

static bool fast_enabled() {                 // PQC_FAST=0: always use the generic sweep kernel
  const char* e = getenv("PQC_FAST");        // (read per launch so tests can compare both paths)
  return !(e && strcmp(e, "0") == 0);
}

static bool seq_enabled() {                  // PQC_SEQ=0: XXZ-type passes stay on k_sweep_pass
  const char* e = getenv("PQC_SEQ");         // (read per launch so tests can compare both paths)
  return !(e && strcmp(e, "0") == 0);
}

static int launch_v1(const V1Args& a_in, cudaStream_t st) {
  V1Args a = a_in;
  a.pf_dist = (a.low_run >= 4) ? prefetch_dist() : 0;
  static PqcDeviceOnce attr_once;
  if (attr_once.first()) {
    PQC_CUDA(cudaFuncSetAttribute(k_sweep_pass<false, false>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    PQC_CUDA(cudaFuncSetAttribute(k_sweep_pass<true, false>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    PQC_CUDA(cudaFuncSetAttribute(k_sweep_pass<false, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    PQC_CUDA(cudaFuncSetAttribute(k_sweep_pass<true, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  }
  const int ipc = 1 << a.items_log2;
  const long long groups = (a.n_items + ipc - 1) / ipc;
  const long long grid = groups << (a.n - a.tb);
  if (grid <= 0) return 0;
  if (grid > 0x7fffffffLL) PQC_FAIL(-1, "pass grid too large; split the batch");
  if (a.hpass && a.hpass->front && (a.npartners != 0 || a.items_log2 != 0 || a.hpass->pipe_idx < 0))
    PQC_FAIL(-5, "internal: front-plan pass outside k_tile_pipe");
  if (a.hpass && a.hpass->pipe_idx >= 0 && a.npartners == 0 && a.nspawn <= TP_MAX_SPAWN &&
      a.items_log2 == 0 &&
      (a.hpass->front ||
       (fast_enabled() && pqc_pipe_enabled() && (a.hpass->fast_ok || seq_enabled())))) {
    PipeArgs f;
    memset(&f, 0, sizeof(f));
    f.src = a.src;
    f.dst = a.dst;
    f.gtrig = a.gtrig;
    f.tstride = a.tstride;
    f.toff = a.toff;
    f.n = a.n;
    memcpy(f.obit, a.obit, sizeof(f.obit));
    f.slots_total = a.slots_total;
    f.active = a.active;
    f.nspawn = a.nspawn;
    for (int k = 0; k < a.nspawn; ++k) {
      f.spawn_slot[k] = a.spawn_slot[k];
      f.spawn_cr[k] = a.hprog->gens[a.spawn_goff[k]].re;
      f.spawn_ci[k] = a.hprog->gens[a.spawn_goff[k]].im;
    }
    f.n_items = a.n_items;
    f.plan = a.hprog->d_pipe + a.hpass->pipe_idx;
    return pqc_pipe_launch(f, a.hprog->h_pipe[a.hpass->pipe_idx], st);
  }
  if (a.hpass && a.hpass->fast_ok && a.npartners == 0 && a.nspawn <= FAST_MAX_SPAWN &&
      fast_enabled()) {
    FastArgs f;
    memset(&f, 0, sizeof(f));
    f.src = a.src;
    f.dst = a.dst;
    f.gtrig = a.gtrig;
    f.tstride = a.tstride;
    f.toff = a.toff;
    f.ntrig = a.ntrig;
    f.wtab = a.wtab;
    f.nwt = a.nwt;
    f.n = a.n;
    memcpy(f.lbit, a.lbit, sizeof(f.lbit));
    memcpy(f.obit, a.obit, sizeof(f.obit));
    f.slots_total = a.slots_total;
    f.active = a.active;
    f.nspawn = a.nspawn;
    for (int k = 0; k < a.nspawn; ++k) {
      f.spawn_slot[k] = a.spawn_slot[k];
      f.spawn_cr[k] = a.hprog->gens[a.spawn_goff[k]].re;
      f.spawn_ci[k] = a.hprog->gens[a.spawn_goff[k]].im;
    }
    f.plan = a.hpass->fast;
    f.pf_dist = a.pf_dist;
    static PqcDeviceOnce fattr_once;
    if (fattr_once.first()) {
      PQC_CUDA(cudaFuncSetAttribute(k_layer_pass<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      PQC_CUDA(cudaFuncSetAttribute(k_layer_pass<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      PQC_CUDA(cudaFuncSetAttribute(k_layer_pass<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      PQC_CUDA(cudaFuncSetAttribute(k_layer_pass<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      PQC_CUDA(cudaFuncSetAttribute(k_layer_pass<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      PQC_CUDA(cudaFuncSetAttribute(k_layer_pass<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    }
    const size_t fsmem = 4096 * sizeof(c128) + (size_t)a.ntrig * sizeof(double2);
    const int hh = pqc_prof_launch_begin((double)a.n_items * 2.0 * sizeof(c128) * (double)(1ll << a.n), st, PQC_PROF_LAYER_PASS);
    const bool g = a.nspawn > 0;
    if (f.plan.ns == 3) {
      if (g) k_layer_pass<3, true><<<(unsigned)grid, 256, fsmem, st>>>(f);
      else k_layer_pass<3, false><<<(unsigned)grid, 256, fsmem, st>>>(f);
    } else if (f.plan.ns == 1) {
      if (g) k_layer_pass<1, true><<<(unsigned)grid, 256, fsmem, st>>>(f);
      else k_layer_pass<1, false><<<(unsigned)grid, 256, fsmem, st>>>(f);
    } else {
      if (g) k_layer_pass<2, true><<<(unsigned)grid, 256, fsmem, st>>>(f);
      else k_layer_pass<2, false><<<(unsigned)grid, 256, fsmem, st>>>(f);
    }
    pqc_prof_launch_end(hh, st);
    PQC_LAUNCH_CHECK();
    return 0;
  }
  if (a.hpass && a.hpass->seq_ok && a.npartners == 0 && a.nspawn <= SEQ_MAX_SPAWN &&
      fast_enabled() && seq_enabled()) {
    SeqArgs f;
    memset(&f, 0, sizeof(f));
    f.src = a.src;
    f.dst = a.dst;
    f.gtrig = a.gtrig;
    f.tstride = a.tstride;
    f.toff = a.toff;
    f.ntrig = a.ntrig;
    f.nwt = a.nwt;
    f.n = a.n;
    memcpy(f.lbit, a.lbit, sizeof(f.lbit));
    memcpy(f.obit, a.obit, sizeof(f.obit));
    f.slots_total = a.slots_total;
    f.active = a.active;
    f.nspawn = a.nspawn;
    for (int k = 0; k < a.nspawn; ++k) {
      f.spawn_slot[k] = a.spawn_slot[k];
      f.spawn_cr[k] = a.hprog->gens[a.spawn_goff[k]].re;
      f.spawn_ci[k] = a.hprog->gens[a.spawn_goff[k]].im;
    }
    f.plan = a.hpass->seq;
    // direct ends need tile positions 0-2 on amplitude bits 0-2 and a first sweep on positions 8-11
    f.staged = (a.hpass->direct_ok && f.plan.geom[0] == 0) ? 0 : 1;
    if (f.staged) {
      // thread bits 0-3 -> the tile positions of amplitude bits 0-3; the other 8 positions in
      // ascending order -> thread bits 4-7, then register bits 0-3
      std::vector<int> rest;
      for (int t = 0; t < 4; ++t) f.st_q[t] = -1;
      for (int pos = 0; pos < 12; ++pos) {
        if (a.lbit[pos] < 4) f.st_q[a.lbit[pos]] = pos;
        else rest.push_back(pos);
      }
      bool okq = rest.size() == 8;
      for (int t = 0; t < 4; ++t) okq = okq && f.st_q[t] >= 0;
      if (!okq) PQC_FAIL(-5, "internal: tile without the four low amplitude bits");
      for (int t = 0; t < 4; ++t) {
        f.st_hi[t] = rest[t];
        const uint32_t i = 1u << rest[4 + t];
        f.st_rs[t] = i ^ (((i >> 3) ^ (i >> 6) ^ (i >> 9)) & 7u);
        f.st_ra[t] = 1u << a.lbit[rest[4 + t]];
      }
    }
    static PqcDeviceOnce qattr_once;
    if (qattr_once.first()) {
#define SEQ_ATTR(G, D, L) PQC_CUDA(cudaFuncSetAttribute(k_layer_seq<G, D, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024))
      SEQ_ATTR(false, false, true); SEQ_ATTR(true, false, true); SEQ_ATTR(false, true, true); SEQ_ATTR(true, true, true);
      SEQ_ATTR(false, false, false); SEQ_ATTR(true, false, false);
#undef SEQ_ATTR
    }
    const size_t qsmem = 4096 * sizeof(c128) + (size_t)a.ntrig * sizeof(double2);
    const int hq = pqc_prof_launch_begin((double)a.n_items * 2.0 * sizeof(c128) * (double)(1ll << a.n), st, PQC_PROF_LAYER_SEQ);
    const bool sp = a.nspawn > 0, dg = f.plan.has_diag != 0;
    bool layers = false;                         // any 4-slot rotation layer op in the pass?
    {
      const SeqPlan& q = a.hpass->seq;
      int tot = 0;
      for (int i = 0; i < q.nsw; ++i) tot = std::max(tot, q.off[i] + q.nops[i]);
      for (int i = 0; i < tot; ++i)
        layers = layers || q.ops[i].kind == PQC_K_LAYER_RX4 || q.ops[i].kind == PQC_K_LAYER_REAL4;
      static const bool force = getenv("PQC_SEQ_LAYERS") != nullptr;      // A/B: always the full op set
      layers = layers || force;
    }
    if (sp && dg) k_layer_seq<true, true, true><<<(unsigned)grid, 256, qsmem, st>>>(f);
    else if (dg) k_layer_seq<false, true, true><<<(unsigned)grid, 256, qsmem, st>>>(f);
    else if (sp && layers) k_layer_seq<true, false, true><<<(unsigned)grid, 256, qsmem, st>>>(f);
    else if (sp) k_layer_seq<true, false, false><<<(unsigned)grid, 256, qsmem, st>>>(f);
    else if (layers) k_layer_seq<false, false, true><<<(unsigned)grid, 256, qsmem, st>>>(f);
    else k_layer_seq<false, false, false><<<(unsigned)grid, 256, qsmem, st>>>(f);
    pqc_prof_launch_end(hq, st);
    PQC_LAUNCH_CHECK();
    return 0;
  }
  static long smem_pad = -1;                 // developer knob: PQC_SMEM_PAD forces 1 CTA per SM
  if (smem_pad < 0) {
    const char* e = getenv("PQC_SMEM_PAD");
    smem_pad = e ? atol(e) : 0;
  }
  const size_t smem = ((size_t)1 << V1_LOCAL_BITS) * sizeof(c128) +
                      (size_t)ipc * a.ntrig * sizeof(double2) + (size_t)smem_pad;
  const int h = pqc_prof_launch_begin((double)a.n_items * 2.0 * sizeof(c128) * (double)(1ll << a.n), st, PQC_PROF_SWEEP_PASS);
  const bool dots = a.npartners > 0, gen = a.nspawn > 0;
  if (dots && gen) k_sweep_pass<true, true><<<(unsigned)grid, V1_NT, smem, st>>>(a);
  else if (dots) k_sweep_pass<true, false><<<(unsigned)grid, V1_NT, smem, st>>>(a);
  else if (gen) k_sweep_pass<false, true><<<(unsigned)grid, V1_NT, smem, st>>>(a);
  else k_sweep_pass<false, false><<<(unsigned)grid, V1_NT, smem, st>>>(a);
  pqc_prof_launch_end(h, st);
  PQC_LAUNCH_CHECK();
  return 0;
}

int pqc_v1_run(const pqc_program* prog, const double* d_angles, long long ld, long long S,
               const c128* d_init, long long init_stride, c128* d_out, cudaStream_t st) {
  const int mode = !d_init ? 1 : (init_stride == 0 ? 2 : 3);
  if (pqc_program_upload(prog)) return -2;
  // sample chunks keep the library-owned trig table below 256 MB
  const long long D = 1ll << prog->n;
  const bool front = pqc_use_front(prog);
  const std::vector<int>& plan = front ? prog->front_run : prog->v1_run;
  const int tj0 = front ? prog->front_tj0 : prog->v1_run_tj0;
  const int ntj = front ? prog->front_ntj : prog->v1_run_ntj;
  const int slots = front ? prog->front_slots : prog->v1_run_slots;
  const long long per = (long long)std::max(1, slots) * (long long)sizeof(double2);
  const long long chunk = std::max<long long>(1, (256ll << 20) / per);
  for (long long c0 = 0; c0 < S; c0 += chunk) {
    const long long c = std::min(chunk, S - c0);
    const double* ang = d_angles ? d_angles + c0 * ld : nullptr;
    c128* out = d_out + c0 * D;
    int rc = trig_prepare(prog, tj0, ntj, slots, ang, ld, c, st);
    if (rc) return rc;
    rc = launch_init(out, mode, mode == 3 ? d_init + c0 * init_stride : d_init, init_stride, c, 1,
                     prog->n, st);
    if (rc) return rc;
    for (int pi : plan) {
      V1Args a;
      memset(&a, 0, sizeof(a));
      fill_pass_args(prog, prog->v1_passes[pi], a);
      a.tstride = slots;
      a.src = out;
      a.dst = out;
      a.n_items = c;
      a.slots_total = 1;
      a.active = 1;
      rc = launch_v1(a, st);
      if (rc) return rc;
    }
  }
  return 0;
}

long long pqc_v1_gpart_elems(const pqc_program* prog, long long S) {
  const long long ntiles = 1ll << std::max(0, prog->n - V1_LOCAL_BITS);
  const long long dots = (long long)(prog->P + 1) * std::max(1, prog->P) * ntiles;
  const long long gram = (long long)gram_ksplit(prog->n) * (prog->P + 1) * (prog->P + 1);
  return S * std::max(dots, gram);
}

int pqc_v1_qfim_reduce(const pqc_program* prog, const c128* d_gpart, long long S, double* d_F,
                       cudaStream_t st) {
  const long long tot = S * prog->P * prog->P;
  if (tot <= 0) return 0;
  const int ntiles = 1 << std::max(0, prog->n - V1_LOCAL_BITS);
  k_qfim_reduce<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(d_gpart, S, prog->P, ntiles, d_F);
  PQC_LAUNCH_CHECK();
  return 0;
}

// index one past the last stage that creates a derivative vector
static size_t last_spawn_stage(const pqc_program* prog) {
  size_t last = 0;
  for (size_t i = 0; i < prog->v1_grad.size(); ++i) {
    const V1Stage& sg = prog->v1_grad[i];
    if (sg.type == 1 || (sg.type == 0 && !prog->v1_passes[sg.pass].spawn_param.empty())) last = i + 1;
  }
  return last;
}

// program ops executed by the passes after the last spawn (pure propagation at the end of the
// derivative plan), ascending
std::vector<int> pqc_v1_trailing_ops(const pqc_program* prog) {
  std::vector<int> ids;
  for (size_t i = last_spawn_stage(prog); i < prog->v1_grad.size(); ++i) {
    const V1Stage& sg = prog->v1_grad[i];
    if (sg.type != 0) continue;
    const V1Pass& ps = prog->v1_passes[sg.pass];
    ids.insert(ids.end(), ps.op_ids.begin(), ps.op_ids.end());
  }
  std::sort(ids.begin(), ids.end());
  return ids;
}

// vector-passes (tile loads + stores of one vector) the QFIM pipeline spends on this plan
long long pqc_v1_plan_cost(const pqc_program* prog) {
  const size_t end = last_spawn_stage(prog);
  long long cost = 0;
  int nlive = 1;
  for (size_t i = 0; i < end; ++i) {
    const V1Stage& sg = prog->v1_grad[i];
    if (sg.type == 0) {
      const V1Pass& ps = prog->v1_passes[sg.pass];
      cost += nlive + (long long)ps.spawn_param.size();
      for (int p : ps.spawn_param) nlive = std::max(nlive, p + 2);
    } else if (sg.type == 1) {
      for (int p : sg.gather_params) nlive = std::max(nlive, p + 2);
      cost += (long long)sg.gather_params.size();
    }
  }
  return cost;
}

int pqc_v1_n_passes(const pqc_program* prog, bool need_final) {
  const size_t end = need_final ? prog->v1_grad.size() : last_spawn_stage(prog);
  int k = 0;
  for (size_t i = 0; i < end; ++i) k += prog->v1_grad[i].type == 0;
  return k;
}

// Forward derivative pipeline on two ping-pong buffers [S][P+1][D].  need_final: also run the
// passes after the last spawn so every vector is the FINAL derivative state (get_gradients);
// without it (QFIM) those unitary passes are skipped because they cannot change an overlap.
// *final_buf receives the buffer holding the vectors after the last executed pass.
int pqc_v1_derivatives(const pqc_program* prog, const double* d_angles, long long ld, long long S,
                       const c128* d_init, long long init_stride, c128* buf_a, c128* buf_b,
                       c128* d_gpart, bool want_dots, bool need_final, c128** final_buf,
                       cudaStream_t st) {
  const int P = prog->P, n = prog->n;
  const int slots_total = P + 1;
  c128* pp[2] = {buf_a, buf_b};
  int cur = 0;                       // buffer holding the current vectors
  int nlive = 1;
  const int mode = !d_init ? 1 : (init_stride == 0 ? 2 : 3);
  const int ntiles = 1 << std::max(0, n - V1_LOCAL_BITS);
  bool first = true;
  if (pqc_program_upload(prog)) return -2;
  {
    int rc = trig_prepare(prog, prog->v1_grad_tj0, prog->v1_grad_ntj, prog->v1_grad_slots, d_angles,
                          ld, S, st);
    if (rc) return rc;
    rc = launch_init(pp[0], mode, d_init, init_stride, S, slots_total, n, st);
    if (rc) return rc;
  }
  const size_t end = need_final ? prog->v1_grad.size() : last_spawn_stage(prog);
  std::vector<int> late_dots;        // Gram columns owed by skipped stages
  for (size_t si = 0; si < prog->v1_grad.size(); ++si) {
    const V1Stage& sg = prog->v1_grad[si];
    if (si >= end) {
      for (int p : sg.partners) late_dots.push_back(p);
      continue;
    }
    if (sg.type == 0) {
      const V1Pass& ps = prog->v1_passes[sg.pass];
      V1Args a;
      memset(&a, 0, sizeof(a));
      fill_pass_args(prog, ps, a);
      a.src = pp[cur];
      a.dst = pp[cur ^ 1];
      a.tstride = prog->v1_grad_slots;
      a.slots_total = slots_total;
      a.active = nlive;
      a.nspawn = (int)ps.spawn_param.size();
      for (int k = 0; k < a.nspawn; ++k) {
        const int p = ps.spawn_param[k];
        a.spawn_slot[k] = 1 + p;
        a.spawn_goff[k] = prog->gen_off[p];
        a.spawn_gcnt[k] = prog->gen_off[p + 1] - prog->gen_off[p];
      }
      a.n_items = S * (a.active + a.nspawn);
      if (want_dots) {
        a.npartners = (int)sg.partners.size();
        for (int q = 0; q < a.npartners; ++q) a.partner_slot[q] = 1 + sg.partners[q];
        a.gpart = d_gpart;
        a.P = P;
        a.ntiles = ntiles;
      }
      const int rc = launch_v1(a, st);
      if (rc) return rc;
      cur ^= 1;
      first = false;
      for (int p : ps.spawn_param) nlive = std::max(nlive, p + 2);
    } else if (sg.type == 1) {
      if (first) PQC_FAIL(-5, "internal: gather before the first pass");
      GatherArgs g;
      memset(&g, 0, sizeof(g));
      g.buf = pp[cur];
      g.n = n;
      g.cb = std::min(n, V1_LOCAL_BITS);
      g.slots_total = slots_total;
      g.gens = prog->d_gens;
      std::vector<int> pauli_params;
      for (int p : sg.gather_params) {
        nlive = std::max(nlive, p + 2);
        if (prog->pspawn[p].type == 1) {
          const int rc = pqc_pair_spawn(pp[cur], n, S, slots_total, 1 + p, prog->pspawn[p],
                                        d_angles, ld, st);
          if (rc) return rc;
        } else {
          pauli_params.push_back(p);
        }
      }
      for (size_t k0 = 0; k0 < pauli_params.size(); k0 += V1_MAX_SPAWN) {
        g.nparams = (int)std::min<size_t>(V1_MAX_SPAWN, pauli_params.size() - k0);
        for (int k = 0; k < g.nparams; ++k) {
          const int p = pauli_params[k0 + k];
          g.slot[k] = 1 + p;
          g.goff[k] = prog->gen_off[p];
          g.gcnt[k] = prog->gen_off[p + 1] - prog->gen_off[p];
          bool pure = g.gcnt[k] > 0;
          for (int t = g.goff[k]; t < g.goff[k] + g.gcnt[k]; ++t) {
            const GenTerm& a = prog->gens[t];
            const GenTerm& a0 = prog->gens[g.goff[k]];
            if (a.zmask || a.re != a0.re || a.im != a0.im) pure = false;
          }
          g.purex[k] = pure ? 1 : 0;
        }
        const long long grid = S << (n - g.cb);
        if (grid > 0x7fffffffLL) PQC_FAIL(-1, "gather grid too large");
        static PqcDeviceOnce gather_attr;
        if (gather_attr.first())
          PQC_CUDA(cudaFuncSetAttribute(k_tile_gather, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        80 * 1024));
        bool all_pure = n >= 12 && g.cb == 12;
        for (int k = 0; k < g.nparams; ++k) all_pure = all_pure && g.purex[k];
        if (all_pure) {
          static PqcDeviceOnce xs_attr;
          if (xs_attr.first()) {
            PQC_CUDA(cudaFuncSetAttribute(k_xsum_gather, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          80 * 1024));
          }
          k_xsum_gather<<<(unsigned)grid, XS_NT, 4096 * sizeof(c128), st>>>(g);
        } else {
          k_tile_gather<<<(unsigned)grid, 256, ((size_t)1 << g.cb) * sizeof(c128), st>>>(g);
        }
        PQC_LAUNCH_CHECK();
      }
    } else {
      for (int p : sg.partners) late_dots.push_back(p);
    }
    // standalone Gram columns (dots-only stages, and whatever skipped stages still owe)
    const bool flush = want_dots && !late_dots.empty() &&
                       (sg.type == 2 || si + 1 == prog->v1_grad.size());
    if (flush) {
      for (size_t q0 = 0; q0 < late_dots.size(); q0 += V1_MAX_PART) {
        V1Args a;
        memset(&a, 0, sizeof(a));
        a.npartners = (int)std::min<size_t>(V1_MAX_PART, late_dots.size() - q0);
        for (int q = 0; q < a.npartners; ++q) a.partner_slot[q] = 1 + late_dots[q0 + q];
        a.gpart = d_gpart;
        a.P = P;
        a.ntiles = ntiles;
        const long long grid = S * nlive * a.npartners;
        if (grid > 0x7fffffffLL) PQC_FAIL(-1, "dot grid too large");
        k_multi_dots<<<(unsigned)grid, 256, 0, st>>>(pp[cur], n, slots_total, nlive, a.npartners, a);
        PQC_LAUNCH_CHECK();
      }
      late_dots.clear();
    }
  }
  if (want_dots && !late_dots.empty()) {
    for (size_t q0 = 0; q0 < late_dots.size(); q0 += V1_MAX_PART) {
      V1Args a;
      memset(&a, 0, sizeof(a));
      a.npartners = (int)std::min<size_t>(V1_MAX_PART, late_dots.size() - q0);
      for (int q = 0; q < a.npartners; ++q) a.partner_slot[q] = 1 + late_dots[q0 + q];
      a.gpart = d_gpart;
      a.P = P;
      a.ntiles = ntiles;
      const long long grid = S * nlive * a.npartners;
      if (grid > 0x7fffffffLL) PQC_FAIL(-1, "dot grid too large");
      k_multi_dots<<<(unsigned)grid, 256, 0, st>>>(pp[cur], n, slots_total, nlive, a.npartners, a);
      PQC_LAUNCH_CHECK();
    }
  }
  if (final_buf) *final_buf = pp[cur];
  return 0;
}
