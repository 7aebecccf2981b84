// ransac_emu.cpp — the batched RANSAC of libkml (kimera-multi_b200/csrc/ransac.cu with geom.cuh
// and fivept_thread.cuh: the source nvcc compiles, launchers included) run on the host under
// tests/emu/cuda_emu.h, in front of the product's own host constants (csrc/sac_host.h).
// Test infrastructure (tests/test_emulated_kernels.py).  The buffer wiring below restates
// run_sac / ransac_batch of lcd.cu (allocation sizes and SacArgs fields).
#include "cuda_emu.h"

// (the CUDA runtime calls the launchers make come from fake_cudart.cpp, linked in)
#include "../../kimera-multi_b200/csrc/ransac.cu"
#include "../../kimera-multi_b200/csrc/sac_host.h"

extern "C" {

// mono = 1: a / b are query / match bearings; mono = 0: query / match 3-D points.  [P][N][3].
// Outputs as kml_ransac_*_batch: models [P][12], n_inliers / iterations / best_draw [P],
// inlier_mask [P][max(1, ceil(N / 32))].  Returns 0, or -5 if the sample stream ran out.
int ransacemu_batch(int mono, int P, int N, const double* a_in, const double* b_in, double threshold, double prob,
                    int max_it, uint32_t seed, int full, int force_generic, double* models, int32_t* n_inliers,
                    int32_t* iterations, int32_t* best_draw, uint32_t* inlier_mask) {
  using namespace kml;
  if (P <= 0) return 0;
  const int S = mono ? 8 : 3;
  const int stride = std::max(N, 8);
  const int mask_words = (stride + 31) / 32;
  std::vector<double> da((size_t)P * stride * 3, 0.0), db((size_t)P * stride * 3, 0.0);
  for (int p = 0; p < P; ++p) {
    memcpy(&da[(size_t)p * stride * 3], a_in + (size_t)p * N * 3, sizeof(double) * 3 * N);
    memcpy(&db[(size_t)p * stride * 3], b_in + (size_t)p * N * 3, sizeof(double) * 3 * N);
  }
  std::vector<int32_t> Ns(P, N);
  std::vector<uint32_t> raw;
  fill_raw_stream(seed, max_it, &raw);
  const int raw_len = (int)raw.size();
  const int cap_draws = sac_cap_draws(raw_len, S, max_it);
  const int n1 = std::max(stride + 1, 512);
  std::vector<double> ktable;
  fill_ktable(n1, S, prob, &ktable);
  const size_t Pa = (size_t)P;
  std::vector<uint16_t> perm(Pa * stride), samples(Pa * kRoundCap * 8);
  std::vector<double> mods(Pa * kRoundCap * 12), fsol, brk, item_q, item_model, best(Pa * 12, 0.0);
  std::vector<int32_t> nroot, valid(Pa * kRoundCap), counts(Pa * kRoundCap), inl(Pa);
  std::vector<uint32_t> fb(4 + 2), item_base, item_list, mask(Pa * mask_words);
  std::vector<uint8_t> item_status;
  std::vector<SacState> st(Pa);
  if (mono) {
    const size_t max_items = Pa * kRoundCap * 10;  // a draw has at most 10 real roots
    nroot.resize(Pa * kRoundCap);
    fsol.resize(Pa * kRoundCap * 130);
    brk.resize(Pa * kRoundCap * 40);
    fb.resize(max_items + 4);
    item_base.resize(Pa * kRoundCap);
    item_list.resize(max_items);
    item_q.resize(max_items);
    item_model.resize(max_items * 12);
    item_status.resize(max_items);
  }
  SacArgs a;
  a.P = P; a.a = da.data(); a.b = db.data(); a.N = Ns.data(); a.stride = stride;
  a.raw = raw.data(); a.raw_len = raw_len; a.cap_draws = cap_draws;
  a.perm = perm.data(); a.samples = samples.data(); a.models = mods.data();
  a.fsol = fsol.data(); a.nroot = nroot.data(); a.brk = brk.data();
  a.fb_list = fb.data() + 4; a.fb_count = fb.data(); a.item_count = fb.data() + 1;
  a.overflow = fb.data() + 2; a.pending = nullptr;
  a.item_cap = (unsigned int)(mono ? Pa * kRoundCap * 10 : 0);
  fb[2] = 0;
  a.item_base = item_base.data(); a.item_list = item_list.data(); a.item_q = item_q.data();
  a.item_model = item_model.data(); a.item_status = item_status.data();
  a.valid = valid.data(); a.counts = counts.data(); a.st = st.data(); a.best_model = best.data();
  a.ktable = ktable.data(); a.ktable_n = n1;
  a.threshold = threshold;
  a.sq_crit = sq_crit_of(threshold);
  a.max_iterations = max_it; a.full = full; a.force_generic = force_generic;
  a.onept = 0; a.prior = nullptr;
  a.first = getenv("KML_EMU_LATENCY_SCHEDULE") ? kSacFirstLatency : kSacFirstThroughput;
  a.n_rounds = getenv("KML_EMU_LATENCY_SCHEDULE") ? kSacRoundsLatency : kSacRoundsThroughput;
  a.alg = (mono && getenv("KML_EMU_STEWENIUS")) ? 1 : 0;
  a.fo_stride = a.alg == 1 ? 130 : 70;
  std::vector<int32_t> active(2 * Pa + 2, 0);
  a.n_active = reinterpret_cast<unsigned int*>(active.data()); a.active = active.data() + 2;
  // per-N sample table (ensure_sample_table of lcd.cu) unless the caller asks for the per-problem sampler
  std::vector<uint16_t> samptab;
  a.sample_tab = nullptr; a.tab_nmax = 0;
  if (stride <= kSampleTabMaxN && !getenv("KML_NO_SAMPLE_TABLE")) {
    samptab.resize((size_t)(stride + 1) * cap_draws * S);
    launch_sample_table(raw.data(), cap_draws, S, stride, samptab.data(), nullptr);
    a.sample_tab = samptab.data(); a.tab_nmax = stride;
  }
  a.inlier_mask = mask.data(); a.mask_words = mask_words; a.n_inliers = inl.data();
  launch_sac_init(a, S, nullptr);
  for (int r = 0; r < a.n_rounds; ++r) {
    if (mono) launch_mono_round(a, r, nullptr); else launch_stereo_round(a, r, nullptr);
  }
  // finish_sac of lcd.cu: rounds are added while a problem's loop has not ended
  a.pending = fb.data() + 3;
  for (int r = a.n_rounds;; ++r) {
    fb[3] = 0;
    launch_sac_pending(a, nullptr);
    if (fb[2]) return -4;  // item lists overflowed (cannot happen at the worst-case size used here)
    if (!fb[3]) break;
    if (mono) launch_mono_round(a, r, nullptr); else launch_stereo_round(a, r, nullptr);
  }
  a.pending = nullptr;
  if (mono) launch_mono_select(a, nullptr); else launch_stereo_select(a, nullptr);
  const int words_out = std::max((N + 31) / 32, 1);
  for (int p = 0; p < P; ++p) {
    if (st[p].exhausted) return -5;
    iterations[p] = st[p].iterations;
    best_draw[p] = st[p].best_draw;
    n_inliers[p] = inl[p];
    memcpy(models + (size_t)p * 12, &best[(size_t)p * 12], 96);
    for (int w = 0; w < words_out; ++w)
      inlier_mask[(size_t)p * words_out + w] = w < (N + 31) / 32 ? mask[(size_t)p * mask_words + w] : 0u;
  }
  return 0;
}

}  // extern "C"
