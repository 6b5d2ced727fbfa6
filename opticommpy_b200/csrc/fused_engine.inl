// Fused four-step engine: step loops (included at the end of ssfm_engine.cu so that it shares the
// plan struct and the launch helpers).  Kernels and the data layout: fused_kernels.cuh.

namespace {

constexpr int kFreqC = 8;  // W positions per k_freq tile (64-byte row segments)

// programmatic dependent launch between the passes of the step loop (OCB_PDL=0 disables; tuning knob)
bool pdl_enabled() {
    const char* v = getenv("OCB_PDL");
    return !(v && atoi(v) == 0);
}

// One-time per-DEVICE setup (function attributes and device limits are per device/context, and one process may
// drive several GPUs): true the first time `slot` is seen on the current device.
enum OnceSlot { ONCE_L2_LIMIT = 0, ONCE_FREQ_BASE = 1, ONCE_TIME_BASE = 16, ONCE_SLOTS = 32 };
bool first_time_on_device(int slot) {
    static bool done[ONCE_SLOTS][64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    bool& d = done[slot][dev & 63];
    if (d) return false;
    d = true;
    return true;
}

template <int Q1, int NP, int MODE>
int launch_time_t(const TimeArgs& a, cudaStream_t st) {
    const int tasks_per_cta = 64 / Q1;
    OCB_LAUNCH_PDL((k_time<Q1, NP, MODE>), a.N2 * NP / tasks_per_cta, 64, 0, st, pdl_enabled(), a);
    return 0;
}
template <int NP, int MODE>
int launch_time_bulk(const TimeArgs& a, cudaStream_t st) {
    if (first_time_on_device(ONCE_TIME_BASE + (NP - 1) * 7 + MODE))
        OCB_CUDA(cudaFuncSetAttribute(k_time_bulk<NP, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      TimeBulkCfg::SMEM_BYTES));
    const int grid = a.N2 < kNumSMs ? a.N2 : kNumSMs;
    OCB_LAUNCH_PDL((k_time_bulk<NP, MODE>), grid, 256, TimeBulkCfg::SMEM_BYTES, st, pdl_enabled(), a);
    return 0;
}
// Time-pass kernel: the one-wave kernel k_time (64-thread CTAs, 7 per SM) by default.  OCB_TIME_KERNEL=bulk (read when a
// plan is created) selects the persistent bulk-copy-fed kernel of fused_time_bulk.cuh for N1 = 1024 (E_c, E_hd and P_ch
// arrive by cp.async.bulk issued before the dependency wait).  Measured on the B200 (profiles/r2_summary.md): launched
// back to back the bulk kernel is faster for the iteration passes (TM_ITERF 21 vs 26 us, TM_ITER 26.6 vs 27.5 us), but
// inside the step loop — alternating with k_freq, cold instruction cache, no overlap of its 193 KB CTAs with the
// neighbouring kernels — the one-wave kernel wins by 2-5 % of the whole step.  Both produce bit-identical fields.
template <int NP, int MODE>
int launch_time(int Q1, const TimeArgs& a, cudaStream_t st) {
    const int choice = a.kernel_choice;
    if (Q1 == 32 && choice == 1) return launch_time_bulk<NP, MODE>(a, st);
    switch (Q1) {
        case 8: return launch_time_t<8, NP, MODE>(a, st);
        case 16: return launch_time_t<16, NP, MODE>(a, st);
        case 32: return launch_time_t<32, NP, MODE>(a, st);
    }
    return fail("fused engine: unsupported N1", __FILE__, __LINE__);
}
// Frequency-pass kernel at N2 = 1024.  Default: k_freq<32, 16> — 128-byte row segments of the W tile by coalesced
// streaming loads / stores, the tile's 128 KB operator slice by cp.async.bulk (TMA engine) before the dependency wait.
// OCB_FREQ_TMA=1 (read when a plan is created): k_freq_tma (fused_freq_tma.cuh) — the tile itself through a CUtensorMap
// (cp.async.bulk.tensor box loads and stores), four independent 128-thread groups per CTA.  Measured on the B200
// (profiles/r2_summary.md): launched back to back both take 15.0 us; inside the step loop the tensor-map variant is 4 %
// slower for the whole step (only half of its operator slice can be prefetched before the wait: 208 KB of shared memory),
// so it is the option, not the default.  All variants produce bit-identical fields.
int freq_c(const ocb_ssfm_plan* p) {
    if (p->q2 != 32) return kFreqC;
    return (p->knob_freq_tma && p->wmap_ok) ? FreqTmaCfg::C : p->knob_freq_c;
}
template <int Q2, int C>
int launch_freq_t(float2* W, const float2* LP, const float2* tw, int N1, int NP, const long long* flag,
                  long long step_id, const long long* need_flag, long long need_id, cudaStream_t st) {
    const size_t smem = FreqCfg<Q2, C>::SMEM_BYTES;
    if (first_time_on_device(ONCE_FREQ_BASE + fft::ilog2(Q2) + (C == 16 ? 6 : 0)))
        OCB_CUDA(cudaFuncSetAttribute(k_freq<Q2, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    OCB_LAUNCH_PDL((k_freq<Q2, C>), NP * N1 / C, Q2 * C, smem, st, pdl_enabled(), W, LP, tw, N1, flag, step_id, need_flag, need_id);
    return 0;
}
int launch_freq(ocb_ssfm_plan* p, const float2* LP, int NP, cudaStream_t st,
                const long long* flag = nullptr, long long step_id = 0, const long long* need_flag = nullptr,
                long long need_id = 0) {
    float2* W = p->G;
    const float2* tw = p->tw2;
    const int N1 = 32 * p->q1, Q2 = p->q2;
    const int c = freq_c(p);
    if (Q2 == 32 && c == FreqTmaCfg::C) {
        const bool lockstep = p->knob_freq_lockstep, st_tma = p->knob_freq_tma_store;
        if (first_time_on_device(ONCE_FREQ_BASE + 14)) {
            OCB_CUDA(cudaFuncSetAttribute(k_freq_tma<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FreqTmaCfg::SMEM_BYTES));
            OCB_CUDA(cudaFuncSetAttribute(k_freq_tma<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FreqTmaCfg::SMEM_BYTES));
            OCB_CUDA(cudaFuncSetAttribute(k_freq_tma<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FreqTmaCfg::SMEM_BYTES));
            OCB_CUDA(cudaFuncSetAttribute(k_freq_tma<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FreqTmaCfg::SMEM_BYTES));
        }
        const int tiles = NP * N1 / FreqTmaCfg::C;
        const int grid = tiles / FreqTmaCfg::GROUPS;  // 128 CTAs (dual-pol): every group of every CTA owns one tile
        OCB_REQUIRE(grid * FreqTmaCfg::GROUPS == tiles && grid <= kNumSMs, "k_freq_tma: unexpected tile count");
        auto kern = lockstep ? (st_tma ? k_freq_tma<true, true> : k_freq_tma<true, false>)
                             : (st_tma ? k_freq_tma<false, true> : k_freq_tma<false, false>);
        OCB_LAUNCH_PDL(kern, grid, 512, FreqTmaCfg::SMEM_BYTES, st, pdl_enabled(), (const CUtensorMap*)p->wmap_dev, W, LP, tw,
                       N1, NP, flag, step_id, need_flag, need_id);
        return 0;
    }
    if (c == 16 && Q2 == 32) return launch_freq_t<32, 16>(W, LP, tw, N1, NP, flag, step_id, need_flag, need_id, st);
    switch (Q2) {
        case 8: return launch_freq_t<8, kFreqC>(W, LP, tw, N1, NP, flag, step_id, need_flag, need_id, st);
        case 16: return launch_freq_t<16, kFreqC>(W, LP, tw, N1, NP, flag, step_id, need_flag, need_id, st);
        case 32: return launch_freq_t<32, kFreqC>(W, LP, tw, N1, NP, flag, step_id, need_flag, need_id, st);
    }
    return fail("fused engine: unsupported N2", __FILE__, __LINE__);
}
int launch_linop_perm(ocb_ssfm_plan* p, float2* LP, double a, double b, double Fs, double h, double scale,
                      cudaStream_t st) {
    OCB_LAUNCH(k_tab_linop_perm, grid_for(p->N, 256, 1), 256, 0, st, LP, p->q1, p->q2,
               freq_c(p), p->N, a, b, Fs, h, scale);
    return 0;
}
// natural planar rows [r][N2*n1 + n2] -> engine layout [r][N1*n2 + n1] (to_engine) or back
int launch_transpose(ocb_ssfm_plan* p, const float2* in, float2* out, int planes, bool to_engine, float scale,
                     cudaStream_t st) {
    const int N1 = 32 * p->q1, N2 = 32 * p->q2;
    const int rows_in = to_engine ? N1 : N2, cols_in = to_engine ? N2 : N1;
    dim3 grid(cols_in / 32, rows_in / 32, planes), block(32, 8);
    OCB_LAUNCH(k_transpose, grid, block, 0, st, in, out, rows_in, cols_in, scale);
    return 0;
}

int fused_init_tables(ocb_ssfm_plan* p, cudaStream_t st) {
    if (p->fused_tables_ready) return 0;
    const int N2 = 32 * p->q2;
    OCB_LAUNCH(k_tab_tw, (32 * p->q1 + 255) / 256, 256, 0, st, p->tw1, p->q1);
    OCB_LAUNCH(k_tab_tw, (32 * p->q2 + 255) / 256, 256, 0, st, p->tw2, p->q2);
    const int64_t items = (int64_t)N2 * 32;
    OCB_LAUNCH(k_tab_inter, (int)((items + 255) / 256), 256, 0, st, p->tabV, p->tabU, N2, p->q1, p->N);
    p->fused_tables_ready = true;
    return 0;
}

TimeArgs time_base(ocb_ssfm_plan* p) {
    TimeArgs a{};
    a.tw = p->tw1; a.tabV = p->tabV; a.tabU = p->tabU;
    a.partials = p->partials; a.sums = p->sums; a.ticket = p->ticket;
    a.N = p->N; a.N2 = 32 * p->q2; a.out_scale = 1.0f; a.cphi = 0.f;
    a.kernel_choice = p->knob_time_kernel;
    return a;
}

}  // namespace

static int fused_manakov_run(ocb_ssfm_plan* p, void* rows_inout, const ocb_manakov_params* q,
                             const void* noise_dev, const int32_t* save_spans, void* save_dev,
                             ocb_manakov_stats* stats, cudaStream_t st) {
    const int64_t N = p->N;
    const int R = 2, K = 1, N1 = 32 * p->q1;
    const double dir = (double)q->direction;
    const double a = -dir * q->alpha_lin / 2.0, b = dir * q->beta2 / 2.0;
    const size_t field_bytes = (size_t)R * N * sizeof(float2);
    if (fused_init_tables(p, st)) return 1;
    // One field buffer: k_time<TM_ITER> replaces the previous iterate by the new one in place, which keeps
    // the per-iteration working set (W, E_c, E_hd, P_ch, operator table = 60 MB at N = 2^20) inside L2.
    float2* bufs[3] = {p->A, p->A, p->A};
    float2* Wb = p->G;
    float2* LP = p->T1;
    int cur = 0;
    ocb_manakov_stats S = {0, 0, 0, 0.0, 0.0};
    double table_h = NAN;
    int next_save = 0;
    uint64_t amp_calls = 0;
    const bool predict = !(getenv("OCB_PREDICT") && atoi(getenv("OCB_PREDICT")) == 0);
    int last_iters = -1;            // iterations of the previous step of this span (-1: unknown)
    bool chained = false;           // first half + iteration 0 of the coming step are already enqueued
    unsigned long long chained_seq = 0;

    // keep the operator tables (read by every k_freq launch) resident in L2 for the duration of this call; the
    // caller's stream gets its previous access-policy window back on every return path
    struct WindowGuard {
        cudaStream_t st; cudaStreamAttrValue old{}; bool have = false;
        ~WindowGuard() { if (have) { cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &old); cudaGetLastError(); } }
    } wguard;
    wguard.st = st;
    wguard.have = cudaStreamGetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &wguard.old) == cudaSuccess;
    cudaGetLastError();
    {
        if (first_time_on_device(ONCE_L2_LIMIT)) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 32u << 20);
        cudaStreamAttrValue attr{};
        attr.accessPolicyWindow.base_ptr = p->T1;
        attr.accessPolicyWindow.num_bytes = (size_t)((char*)p->Pch - (char*)p->T1);  // T1 and T2
        attr.accessPolicyWindow.hitRatio = 1.0f;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);  // best effort
        cudaGetLastError();
    }
    if (launch_transpose(p, (const float2*)rows_inout, bufs[0], R, true, 1.0f, st)) return 1;
    const float2* noiseT = nullptr;
    if (q->direction > 0 && q->amp_mode == OCB_AMP_EDFA && q->noise_mode == OCB_NOISE_INJECTED) {
        if (launch_transpose(p, (const float2*)noise_dev, p->Nb, K, true, 1.0f, st)) return 1;
        noiseT = p->Nb;
    }

    for (int span = 1; span <= q->n_spans; ++span) {
        if (q->direction < 0 && q->amp_mode != OCB_AMP_NONE)
            if (launch_amp(bufs[cur], R, N, exp(-q->alpha_lin / 2.0 * q->Lspan), 0.0, nullptr, 1, 0, 0, st)) return 1;
        double maxP = 0.0;
        if (q->nlprMethod) {
            if (launch_power_stats(p, bufs[cur], N, K, st)) return 1;
            if (fetch_sums(p, st)) return 1;
            maxP = p->h_sums[2];
        }
        double z = 0.0;
        last_iters = -1;  // the amplifier changed the power: no prediction for the first step of a span
        while (z < q->Lspan) {  // channels.py:387
            double hz_;
            if (q->nlprMethod) {  // channels.py:392-397
                const double cand = q->maxNlinPhaseRot / ((8.0 / 9.0) * q->gamma * maxP);
                hz_ = (q->Lspan - z >= cand) ? cand : (q->Lspan - z);
            } else if (q->Lspan - z < q->hz) {  // channels.py:398-401
                hz_ = q->Lspan - z;
            } else {
                hz_ = q->hz;
            }
            if (!(table_h == hz_)) {
                if (launch_linop_perm(p, LP, a, b, q->Fs, hz_ / 2.0, 1.0 / (double)N, st)) return 1;
                table_h = hz_;
            }
            // ---- one SSFM step --------------------------------------------------------------------------
            // Launch protocol (no stream synchronisation anywhere):
            //  * iteration it+1 is enqueued BEFORE the outcome of iteration it is known; the finalising block
            //    of `it` sets a device flag when lim < tol and the speculative launches of the same step exit.
            //  * PREDICTION: in a fixed-step span the iteration count hardly ever changes from one step to the
            //    next.  The iteration predicted to be the last one (index `pred`) runs as TM_ITERF, which
            //    transforms the new iterate itself, and the first half (+ iteration 0) of the NEXT step is
            //    enqueued behind it, guarded by a second flag that TM_ITERF sets only if it did converge.  A
            //    right prediction saves the TM_FWD pass and the E_hd/P_ch traffic of the last iteration; a wrong
            //    one costs a few empty launches plus the TM_ROT recovery pass.  Decisions (lim < tol) and
            //    arithmetic are those of the plain flow, so step/iteration counts and fields are unchanged.
            const long long step_id = ++p->step_counter;
            const bool speculate = !p->prof_on;
            const float cphi_first = (float)(dir * hz_ * (8.0 / 9.0) * q->gamma);
            const float cphi_iter = (float)(dir * hz_ * (8.0 / 9.0) * q->gamma * 0.5);
            auto time_args_iter = [&](long long sid, unsigned long long seq, bool final_pred, const long long* need,
                                      long long need_id) {
                TimeArgs ci = time_base(p);
                ci.in = Wb; ci.out = Wb; ci.aux0 = p->A; ci.aux1 = p->A; ci.ehd = p->Ehd; ci.pch = p->Pch;
                ci.cphi = cphi_iter;
                ci.ext.mail = p->d_mail; ci.ext.converged_step = p->conv_flag; ci.ext.step_id = sid;
                ci.ext.seq = seq; ci.ext.tol = q->tol;
                ci.ext.final_step = final_pred ? p->final_flag : nullptr;
                ci.need_flag = need; ci.need_id = need_id;
                return ci;
            };
            // k_freq + time pass of one fixed-point iteration (channels.py:420-421, 424, 436, 414-417)
            auto enqueue_iter = [&](long long sid, unsigned long long seq, bool final_pred, const long long* need,
                                    long long need_id) -> int {
                {
                    ProfScope ps(p, 2, st);
                    if (launch_freq(p, LP, R, st, p->conv_flag, sid, need, need_id)) return 1;
                }
                ProfScope ps(p, 0, st);
                const TimeArgs ci = time_args_iter(sid, seq, final_pred, need, need_id);
                return final_pred ? launch_time<2, TM_ITERF>(p->q1, ci, st) : launch_time<2, TM_ITER>(p->q1, ci, st);
            };
            // first half step from the frequency-side buffer (channels.py:409-410 after the forward time pass):
            // frequency pass with L, then the time pass that completes the inverse, stores E_hd / Pch and rotates
            auto enqueue_first_half = [&](bool with_fwd, const long long* need, long long need_id) -> int {
                {
                    ProfScope ps(p, 2, st);
                    if (with_fwd) {
                        TimeArgs c0 = time_base(p);
                        c0.in = p->A; c0.out = Wb;
                        if (launch_time<2, TM_FWD>(p->q1, c0, st)) return 1;
                    }
                    if (launch_freq(p, LP, R, st, nullptr, 0, need, need_id)) return 1;
                }
                ProfScope ps(p, 1, st);
                TimeArgs c1 = time_base(p);
                c1.in = Wb; c1.out = Wb; c1.aux0 = p->A; c1.aux1 = p->Ehd; c1.pch = p->Pch;
                c1.cphi = cphi_first;
                c1.need_flag = need; c1.need_id = need_id;
                return launch_time<2, TM_FIRST>(p->q1, c1, st);
            };

            const bool can_predict = predict && speculate && !q->nlprMethod;
            int pred = (can_predict && last_iters >= 1 && last_iters <= q->maxIter) ? last_iters - 1 : -1;
            unsigned long long seq_cur;
            if (chained) {  // first half and iteration 0 of this step were enqueued behind the previous TM_ITERF
                seq_cur = chained_seq;
                chained = false;
            } else {
                if (enqueue_first_half(true, nullptr, 0)) return 1;
                seq_cur = ++p->mail_seq;
                if (enqueue_iter(step_id, seq_cur, pred == 0, nullptr, 0)) return 1;
            }
            int iters_this_step = 0;
            for (int it = 0; it < q->maxIter; ++it) {
                unsigned long long seq_next = 0;
                bool next_enqueued = false;
                if (it == pred) {
                    // TM_ITERF is in flight: chain the next step behind it if that step uses the same operator table
                    double zn = z + hz_;
                    const bool next_same = (zn < q->Lspan) && !(q->Lspan - zn < q->hz) && (q->hz == hz_);
                    if (next_same) {
                        const long long next_id = p->step_counter + 1;
                        if (enqueue_first_half(false, p->final_flag, step_id)) return 1;
                        chained_seq = ++p->mail_seq;
                        if (enqueue_iter(next_id, chained_seq, pred == 0, p->final_flag, step_id)) return 1;
                        chained = true;
                    }
                } else if (speculate && it + 1 < q->maxIter) {
                    seq_next = ++p->mail_seq;
                    if (enqueue_iter(step_id, seq_next, it + 1 == pred, nullptr, 0)) return 1;
                    next_enqueued = true;
                }
                if (wait_mail(p, seq_cur, st)) return 1;
                const double lim = sqrt(p->h_sums[0]) / sqrt(p->h_sums[1]);  // channels.py:517-519
                const bool converged = p->h_sums[3] != 0.0;  // the device's decision (same formula, same doubles)
                S.iterations++;
                iters_this_step++;
                S.last_lim = lim;
                if (converged) {  // channels.py:429
                    if (it != pred) chained = false;
                    break;
                }
                if (it == pred) {
                    // wrong prediction: the chained launches exit on the flag; redo the second half of a plain
                    // iteration (rotation of E_hd with the new phase + forward time pass)
                    chained = false;
                    pred = -1;
                    if (it < q->maxIter - 1) {
                        TimeArgs cr = time_base(p);
                        cr.out = Wb; cr.aux0 = p->A; cr.ehd = p->Ehd; cr.pch = p->Pch; cr.cphi = cphi_iter;
                        if (launch_time<2, TM_ROT>(p->q1, cr, st)) return 1;
                    }
                }
                if (it == q->maxIter - 1) { S.nonconverged++; break; }  // channels.py:431-434
                if (!next_enqueued) {
                    seq_next = ++p->mail_seq;
                    if (enqueue_iter(step_id, seq_next, it + 1 == pred, nullptr, 0)) return 1;
                }
                seq_cur = seq_next;
            }
            last_iters = iters_this_step;
            maxP = p->h_sums[2];
            z += hz_;
            S.steps++;
            S.z_last_step = hz_;
        }
        if (q->direction > 0) {  // channels.py:443-451
            if (q->amp_mode == OCB_AMP_EDFA) {
                if (launch_amp(bufs[cur], R, N, sqrt(q->edfa_gain_lin), noiseT ? 0.0 : sqrt(q->edfa_noise_var / 2.0),
                               noiseT, K, q->seed, amp_calls++, st, N1, 32 * p->q2)) return 1;
            } else if (q->amp_mode == OCB_AMP_IDEAL) {
                if (launch_amp(bufs[cur], R, N, exp(q->alpha_lin / 2.0 * q->Lspan), 0.0, nullptr, 1, 0, 0, st)) return 1;
            }
        }
        if (next_save < q->n_save && save_spans[next_save] == span) {  // channels.py:453-456
            if (launch_transpose(p, bufs[cur], (float2*)((char*)save_dev + (size_t)next_save * field_bytes), R, false,
                                 1.0f, st)) return 1;
            next_save++;
        }
    }
    if (launch_transpose(p, bufs[cur], (float2*)rows_inout, R, false, 1.0f, st)) return 1;
    if (stats) *stats = S;
    return 0;
}

static int fused_nlse_run(ocb_ssfm_plan* p, void* row_inout, const ocb_nlse_params* q, const void* noise_dev,
                          cudaStream_t st) {
    const int64_t N = p->N;
    const int N1 = 32 * p->q1;
    float2* E = p->A;  // engine-layout copy of the field
    float2* Wb = p->G;
    const double a = -q->alpha_lin / 2.0, b = q->beta2 / 2.0;
    if (fused_init_tables(p, st)) return 1;
    // LP1 = L/N (entering / leaving a span), LP2 = L^2/N (between two steps): channels.py:221 + :229 merged
    if (launch_linop_perm(p, p->T1, a, b, q->Fs, q->hz / 2.0, 1.0 / (double)N, st)) return 1;
    if (launch_linop_perm(p, p->T2, a, b, q->Fs, q->hz, 1.0 / (double)N, st)) return 1;
    if (launch_transpose(p, (const float2*)row_inout, E, 1, true, 1.0f, st)) return 1;
    const bool inj = (q->amp_mode == OCB_AMP_EDFA && q->noise_mode == OCB_NOISE_INJECTED);
    const float2* noiseT = nullptr;
    if (inj) {
        if (launch_transpose(p, (const float2*)noise_dev, p->Nb, 1, true, 1.0f, st)) return 1;
        noiseT = p->Nb;
    }
    for (int span = 0; span < q->n_spans; ++span) {
        float gain = 1.0f;
        if (q->amp_mode == OCB_AMP_IDEAL) gain = (float)exp(q->alpha_lin / 2.0 * q->n_steps * q->hz);
        if (q->amp_mode == OCB_AMP_EDFA) gain = (float)sqrt(q->edfa_gain_lin);
        if (q->n_steps > 0) {
            TimeArgs cf = time_base(p);
            cf.in = E; cf.out = Wb;
            if (launch_time<1, TM_FWD>(p->q1, cf, st)) return 1;  // channels.py:216 (time half)
            for (int s = 0; s < q->n_steps; ++s) {
                if (launch_freq(p, s == 0 ? p->T1 : p->T2, 1, st)) return 1;
                TimeArgs cn = time_base(p);
                cn.in = Wb; cn.out = Wb; cn.cphi = (float)(q->gamma * q->hz);
                if (launch_time<1, TM_NLSE>(p->q1, cn, st)) return 1;  // :224-228
            }
            if (launch_freq(p, p->T1, 1, st)) return 1;  // :229 of the last step
            TimeArgs ci = time_base(p);
            ci.in = Wb; ci.out = E; ci.out_scale = gain;
            if (launch_time<1, TM_INV>(p->q1, ci, st)) return 1;  // :232 + gain of :234/:236
            if (q->amp_mode == OCB_AMP_EDFA)  // noise only; the gain rode in the inverse time pass
                if (launch_amp(E, 1, N, 1.0, inj ? 0.0 : sqrt(q->edfa_noise_var / 2.0), noiseT, 1, q->seed,
                               (uint64_t)span, st, N1, 32 * p->q2)) return 1;
        } else if (q->amp_mode != OCB_AMP_NONE) {
            if (launch_amp(E, 1, N, gain, (q->amp_mode == OCB_AMP_EDFA && !inj) ? sqrt(q->edfa_noise_var / 2.0) : 0.0,
                           noiseT, 1, q->seed, (uint64_t)span, st, N1, 32 * p->q2)) return 1;
        }
    }
    if (launch_transpose(p, E, (float2*)row_inout, 1, false, 1.0f, st)) return 1;
    return 0;
}

// Measurement aid: see include/opticomm_b200.h (ocb_ssfm_plan_pass_time)
extern "C" int ocb_ssfm_plan_pass_time(ocb_ssfm_plan* p, int which, int reps, double* avg_us, void* stream) {
    OCB_REQUIRE(p && avg_us && reps > 0, "pass_time: bad argument");
    OCB_REQUIRE(p->ws != nullptr && p->fused_ok && p->rows == 2, "pass_time: needs a bound dual-pol plan of the fused engine");
    cudaStream_t st = (cudaStream_t)stream;
    if (fused_init_tables(p, st)) return 1;
    const int N1 = 32 * p->q1;
    if (launch_linop_perm(p, p->T1, 0.0, 0.0, 1.0, 1.0, 1.0 / (double)p->N, st)) return 1;  // unit operator
    cudaEvent_t e0, e1;
    OCB_CUDA(cudaEventCreate(&e0));
    OCB_CUDA(cudaEventCreate(&e1));
    auto one = [&]() -> int {
        TimeArgs a = time_base(p);
        a.in = p->G; a.out = p->G; a.aux0 = p->A; a.aux1 = p->A; a.ehd = p->Ehd; a.pch = p->Pch; a.cphi = 1e-3f;
        switch (which) {
            case 0: return launch_freq(p, p->T1, 2, st);
            case 1: a.aux1 = p->Ehd; return launch_time<2, TM_FIRST>(p->q1, a, st);
            case 2: return launch_time<2, TM_ITER>(p->q1, a, st);
            case 3: return launch_time<2, TM_ITERF>(p->q1, a, st);
            case 4: a.in = p->A; return launch_time<2, TM_FWD>(p->q1, a, st);
        }
        return fail("pass_time: unknown pass", __FILE__, __LINE__);
    };
    for (int i = 0; i < 3; ++i) if (one()) return 1;  // warm-up
    OCB_CUDA(cudaEventRecord(e0, st));
    for (int i = 0; i < reps; ++i) if (one()) return 1;
    OCB_CUDA(cudaEventRecord(e1, st));
    OCB_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    OCB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *avg_us = 1e3 * (double)ms / reps;
    return 0;
}
