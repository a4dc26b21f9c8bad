// TEST-ONLY host instantiation of bayesnetworks_b200/csrc/*_core.cuh with a
// one-lane "warp".  It exists so that the sequential chain logic (draw order,
// stale members, speculative windows, ancestor bitsets, trace rows) can be
// checked against the oracle on a machine without a GPU.  It is built by
// tests/ into tests/emu/_build/ and is never part of libbn_b200.so.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "chain_core.cuh"

using namespace bn;

extern "C" int emu_run_chain(
    int P, int max_par, int n_samples, const double* C, const unsigned char* node_type,
    const unsigned char* sim_edge, int n_sim_edges, double phi, double omega,
    int initial_network, int drop, int n_iter, int output_every,
    const int* prior_par, const int* prior_npar,
    int rng_kind, const int* seeds, const unsigned int* mt_state, const double* replay,
    long replay_len, int capacity,
    int* t_iter, int* t_changed, int* t_movetype, double* t_gll, int* t_add, int* t_del,
    int* t_fn, int* t_fp, int moves_capacity, int* moves, int* edge_freq, int* npar_freq,
    int* final_par, int* final_npar, long* out_counters /*[12]*/) {
  ChainParams p;
  p.P = P; p.max_par = max_par; p.W = (P + 31) / 32; p.Ws = (p.W + 3) / 4 * 4; p.n_samples = n_samples;
  std::vector<double> diag(P);
  for (int i = 0; i < P; i++) diag[i] = C[(size_t)i * P + i];
  p.C = C; p.ldc = P; p.diag = diag.data(); p.node_type = node_type; p.sim_edge = sim_edge;
  p.n_sim_edges = n_sim_edges; p.phi = phi; p.omega = omega;
  p.initial_network = initial_network; p.drop = drop; p.n_iter = n_iter;
  p.output_every = output_every; p.trace_capacity = capacity; p.moves_capacity = moves_capacity;
  p.prior_par = prior_par; p.prior_npar = prior_npar; p.prior_stride = max_par;
  set_row_geom(p);
  std::vector<double> ratio(max_par + 10);
  for (int k = 0; k < (int)ratio.size(); k++) ratio[k] = (double)(n_samples - 1) / (double)(n_samples - k - 1);
  p.sc.half_n = (double)n_samples / 2.0;
  p.sc.ratio = ratio.data();

  std::vector<int> par((size_t)P * max_par), npar(P), born((size_t)P * max_par);
  std::vector<int> scratch((size_t)scratch_words(P, 1, 1)), hp_list(P);
  std::vector<double> base(P);
  std::vector<uint32_t> anc_store((size_t)P * p.Ws + 4), haspar(p.W);
  uint32_t* anc = (uint32_t*)(((uintptr_t)anc_store.data() + 15) & ~(uintptr_t)15);
  std::vector<uint32_t> mt(624);
  if (mt_state) memcpy(mt.data(), mt_state, 624 * 4);
  ChainMem m;
  m.par = par.data(); m.npar = npar.data(); m.born = born.data(); m.base = base.data();
  m.anc = anc; m.haspar = haspar.data();
  m.scratch = scratch.data(); m.hp_list = hp_list.data(); m.helper = nullptr;
  m.t_iter = t_iter; m.t_changed = t_changed; m.t_movetype = t_movetype; m.t_gll = t_gll;
  m.t_add = t_add; m.t_del = t_del; m.t_fn = t_fn; m.t_fp = t_fp;
  m.moves = moves; m.edge_freq = edge_freq;
  std::vector<int> npar_since(P);
  m.npar_freq = npar_freq; m.npar_since = npar_since.data();
  std::vector<double> dscore((size_t)P * max_par, NAN);
  m.dscore = dscore.data();
  std::vector<double> fac((size_t)P * fac_stride(fac_mp(max_par)) + 2), rowbuf((size_t)REPLAY_POS * row_stride(fac_mp(max_par)) + 2);
  m.fac = (double*)(((uintptr_t)fac.data() + 15) & ~(uintptr_t)15);
  m.rowbuf = (double*)(((uintptr_t)rowbuf.data() + 15) & ~(uintptr_t)15);

  std::vector<double> ubuf(RNG_CAP);
  RngStream rng;
  if (rng_kind == RNG_WH) rng_init_wh(rng, seeds[0], seeds[1], seeds[2], ubuf.data());
  else if (rng_kind == RNG_RMT) rng_init_rmt(rng, mt.data(), 624, ubuf.data());
  else rng_init_replay(rng, replay, replay_len, ubuf.data());

  ChainScalars s;
  WindowSlots ws;
  if (max_par <= 8) run_chain<8>(p, m, s, rng, ws);
  else if (max_par <= 16) run_chain<16>(p, m, s, rng, ws);
  else if (max_par <= 64) run_chain<64>(p, m, s, rng, ws);
  else return 7;

  memcpy(final_par, par.data(), sizeof(int) * par.size());
  memcpy(final_npar, npar.data(), sizeof(int) * P);
  out_counters[0] = s.read_pos; out_counters[1] = s.valid_iters;
  for (int t = 0; t < 3; t++) { out_counters[2 + t] = s.proposed[t]; out_counters[5 + t] = s.reject[t]; }
  out_counters[8] = s.n_rows; out_counters[9] = s.n_moves; out_counters[10] = s.n_nonpd;
  out_counters[11] = s.windows;
  return s.status;
}

#if defined(BN_EMU_STATS)
extern "C" void emu_get_stats(long* out) {
  const EmuStats& s = emu_stats();
  const long v[9] = {s.del_total, s.del_trivial, s.del_desc, s.del_rounds, s.del_rows_eval, s.del_lost_bits,
                     s.add_total, s.add_trivial, s.add_desc};
  for (int i = 0; i < 9; i++) out[i] = v[i];
}
#endif

extern "C" double emu_score_set(const double* C, int P, int c, const int* S, int k, int n_samples) {
  std::vector<double> L((size_t)k * (k + 1) / 2 + 1), z(k + 1);
  int npd = 0;
  return score_set(C, P, c, S, k, n_samples, L.data(), z.data(), &npd);
}

extern "C" void emu_uniforms(int rng_kind, const int* seeds, const unsigned int* mt_state, int n,
                             double* out) {
  std::vector<double> ubuf(RNG_CAP);
  std::vector<uint32_t> mt(624);
  if (mt_state) memcpy(mt.data(), mt_state, 624 * 4);
  RngStream rng;
  if (rng_kind == RNG_WH) rng_init_wh(rng, seeds[0], seeds[1], seeds[2], ubuf.data());
  else rng_init_rmt(rng, mt.data(), 624, ubuf.data());
  for (int i = 0; i < n; i++) {
    while (rng.gen_hi <= i) rng_fill_chunk(rng);
    out[i] = ubuf[i & (RNG_CAP - 1)];
  }
}
