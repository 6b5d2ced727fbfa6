// Fused four-step engine: plan extension + step loops (included at the end of ssfm_engine.cu so
// that it shares the plan struct and the launch helpers).  Kernels: fused_kernels.cuh.

namespace {

constexpr int kColC_Manakov = 4;  // columns per tile (both pols in one CTA -> 256 threads at Q1 = 32)
constexpr int kColC_Nlse = 8;     // single pol -> 256 threads at Q1 = 32

size_t col_smem(int Q1, int C, int NP) {
    return (size_t)32 * Q1 * 8 + (size_t)C * 32 * 8 + (size_t)C * Q1 * 8 + (size_t)NP * 2 * 32 * (Q1 * C + C) * 4;
}
size_t row_smem(int Q2) {
    const int gbuf = 32 * (Q2 + 1) + (Q2 < 32 ? Q2 : 0);
    return (size_t)32 * Q2 * 8 + (size_t)(256 / Q2) * 2 * gbuf * 4;
}

template <int Q1, int C, int NP, int MODE>
int launch_col_t(const ColArgs& a, cudaStream_t st) {
    static bool configured = false;
    const size_t smem = col_smem(Q1, C, NP);
    if (!configured) {
        OCB_CUDA(cudaFuncSetAttribute(k_col<Q1, C, NP, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    OCB_LAUNCH((k_col<Q1, C, NP, MODE>), a.N2 / C, NP * Q1 * C, smem, st, a);
    return 0;
}
template <int C, int NP, int MODE>
int launch_col(int Q1, const ColArgs& a, cudaStream_t st) {
    switch (Q1) {
        case 8: return launch_col_t<8, C, NP, MODE>(a, st);
        case 16: return launch_col_t<16, C, NP, MODE>(a, st);
        case 32: return launch_col_t<32, C, NP, MODE>(a, st);
    }
    return fail("fused engine: unsupported N1", __FILE__, __LINE__);
}
template <int Q2>
int launch_row_t(float2* W, const float2* LP, const float2* tw, int N1, int64_t n_rows, cudaStream_t st) {
    static int grid_cap = 0;
    const size_t smem = row_smem(Q2);
    if (!grid_cap) {
        OCB_CUDA(cudaFuncSetAttribute(k_row<Q2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        OCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_row<Q2>, 256, smem));
        grid_cap = kNumSMs * (per_sm > 0 ? per_sm : 1);
    }
    const int64_t blocks = n_rows / (256 / Q2);
    const int grid = (int)(blocks < grid_cap ? blocks : grid_cap);
    OCB_LAUNCH(k_row<Q2>, grid, 256, smem, st, W, LP, tw, N1, n_rows);
    return 0;
}
int launch_row(int Q2, float2* W, const float2* LP, const float2* tw, int N1, int64_t n_rows, cudaStream_t st) {
    switch (Q2) {
        case 8: return launch_row_t<8>(W, LP, tw, N1, n_rows, st);
        case 16: return launch_row_t<16>(W, LP, tw, N1, n_rows, st);
        case 32: return launch_row_t<32>(W, LP, tw, N1, n_rows, st);
    }
    return fail("fused engine: unsupported N2", __FILE__, __LINE__);
}
int launch_linop_perm(ocb_ssfm_plan* p, float2* LP, double a, double b, double Fs, double h, double scale,
                      cudaStream_t st) {
    const int g = grid_for(p->N, 256, 1);
    const int N1 = 32 * p->q1;
    switch (p->q2) {
        case 8: OCB_LAUNCH(k_tab_linop_perm<8>, g, 256, 0, st, LP, N1, p->N, a, b, Fs, h, scale); break;
        case 16: OCB_LAUNCH(k_tab_linop_perm<16>, g, 256, 0, st, LP, N1, p->N, a, b, Fs, h, scale); break;
        case 32: OCB_LAUNCH(k_tab_linop_perm<32>, g, 256, 0, st, LP, N1, p->N, a, b, Fs, h, scale); break;
        default: return fail("fused engine: unsupported N2", __FILE__, __LINE__);
    }
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

ColArgs col_base(ocb_ssfm_plan* p) {
    ColArgs a{};
    a.tw = p->tw1; a.tabV = p->tabV; a.tabU = p->tabU;
    a.partials = p->partials; a.sums = p->sums; a.ticket = p->ticket;
    a.N = p->N; a.N2 = 32 * p->q2; a.out_scale = 1.0f; a.cphi = 0.f;
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
    float2* bufs[3] = {(float2*)rows_inout, p->A, p->B};
    float2* Wb = p->G;
    float2* LP = p->T1;
    int cur = 0;
    ocb_manakov_stats S = {0, 0, 0, 0.0, 0.0};
    double table_h = NAN;
    int next_save = 0;
    uint64_t amp_calls = 0;
    constexpr int C = kColC_Manakov;

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
        while (z < q->Lspan) {
            double hz_;
            if (q->nlprMethod) {
                const double cand = q->maxNlinPhaseRot / ((8.0 / 9.0) * q->gamma * maxP);
                hz_ = (q->Lspan - z >= cand) ? cand : (q->Lspan - z);
            } else if (q->Lspan - z < q->hz) {
                hz_ = q->Lspan - z;
            } else {
                hz_ = q->hz;
            }
            if (!(table_h == hz_)) {
                if (launch_linop_perm(p, LP, a, b, q->Fs, hz_ / 2.0, 1.0 / (double)N, st)) return 1;
                table_h = hz_;
            }
            // first half step: column FFT of the step-start field, row pass with L, then the column
            // kernel that finishes the inverse, stores E_hd / Pch, rotates and goes forward again
            {
                ProfScope ps(p, 2, st);
                ColArgs c0 = col_base(p);
                c0.in = bufs[cur]; c0.out = Wb;
                if (launch_col<C, 2, COL_FWD>(p->q1, c0, st)) return 1;
                if (launch_row(p->q2, Wb, LP, p->tw2, N1, (int64_t)R * N1, st)) return 1;
            }
            {
                ProfScope ps(p, 1, st);
                ColArgs c1 = col_base(p);
                c1.in = Wb; c1.out = Wb; c1.aux0 = bufs[cur]; c1.aux1 = p->Ehd; c1.pch = p->Pch;
                c1.cphi = (float)(dir * hz_ * (8.0 / 9.0) * q->gamma);
                if (launch_col<C, 2, COL_FIRST>(p->q1, c1, st)) return 1;
            }
            int ec = cur, dst = (cur + 1) % 3;
            for (int it = 0; it < q->maxIter; ++it) {
                {
                    ProfScope ps(p, 2, st);
                    if (launch_row(p->q2, Wb, LP, p->tw2, N1, (int64_t)R * N1, st)) return 1;
                }
                {
                    ProfScope ps(p, 0, st);
                    ColArgs ci = col_base(p);
                    ci.in = Wb; ci.out = Wb; ci.aux0 = bufs[ec]; ci.aux1 = bufs[dst]; ci.ehd = p->Ehd; ci.pch = p->Pch;
                    ci.cphi = (float)(dir * hz_ * (8.0 / 9.0) * q->gamma * 0.5);
                    if (launch_col<C, 2, COL_ITER>(p->q1, ci, st)) return 1;
                }
                if (fetch_sums(p, st)) return 1;
                const double lim = sqrt(p->h_sums[0]) / sqrt(p->h_sums[1]);
                S.iterations++;
                S.last_lim = lim;
                const int third = 3 - ec - dst;
                ec = dst;
                dst = third;
                if (lim < q->tol) break;
                if (it == q->maxIter - 1) S.nonconverged++;
            }
            cur = ec;
            maxP = p->h_sums[2];
            z += hz_;
            S.steps++;
            S.z_last_step = hz_;
        }
        if (q->direction > 0) {
            if (q->amp_mode == OCB_AMP_EDFA) {
                const bool inj = (q->noise_mode == OCB_NOISE_INJECTED);
                if (launch_amp(bufs[cur], R, N, sqrt(q->edfa_gain_lin), inj ? 0.0 : sqrt(q->edfa_noise_var / 2.0),
                               inj ? (const float2*)noise_dev : nullptr, K, q->seed, amp_calls++, st)) return 1;
            } else if (q->amp_mode == OCB_AMP_IDEAL) {
                if (launch_amp(bufs[cur], R, N, exp(q->alpha_lin / 2.0 * q->Lspan), 0.0, nullptr, 1, 0, 0, st)) return 1;
            }
        }
        if (next_save < q->n_save && save_spans[next_save] == span) {
            OCB_CUDA(cudaMemcpyAsync((char*)save_dev + (size_t)next_save * field_bytes, bufs[cur], field_bytes,
                                     cudaMemcpyDeviceToDevice, st));
            next_save++;
        }
    }
    if (cur != 0) OCB_CUDA(cudaMemcpyAsync(rows_inout, bufs[cur], field_bytes, cudaMemcpyDeviceToDevice, st));
    if (stats) *stats = S;
    return 0;
}

static int fused_nlse_run(ocb_ssfm_plan* p, void* row_inout, const ocb_nlse_params* q, const void* noise_dev,
                          cudaStream_t st) {
    const int64_t N = p->N;
    const int N1 = 32 * p->q1;
    float2* E = (float2*)row_inout;
    float2* Wb = p->G;
    const double a = -q->alpha_lin / 2.0, b = q->beta2 / 2.0;
    if (fused_init_tables(p, st)) return 1;
    // LP1 = L/N (entering / leaving a span), LP2 = L^2/N (between two steps): channels.py:221 + :229 merged
    if (launch_linop_perm(p, p->T1, a, b, q->Fs, q->hz / 2.0, 1.0 / (double)N, st)) return 1;
    if (launch_linop_perm(p, p->T2, a, b, q->Fs, q->hz, 1.0 / (double)N, st)) return 1;
    constexpr int C = kColC_Nlse;
    for (int span = 0; span < q->n_spans; ++span) {
        float gain = 1.0f;
        if (q->amp_mode == OCB_AMP_IDEAL) gain = (float)exp(q->alpha_lin / 2.0 * q->n_steps * q->hz);
        const bool inj = (q->noise_mode == OCB_NOISE_INJECTED);
        if (q->n_steps > 0) {
            ColArgs cf = col_base(p);
            cf.in = E; cf.out = Wb;
            if (launch_col<C, 1, COL_FWD>(p->q1, cf, st)) return 1;  // channels.py:216 (column half)
            for (int s = 0; s < q->n_steps; ++s) {
                if (launch_row(p->q2, Wb, s == 0 ? p->T1 : p->T2, p->tw2, N1, N1, st)) return 1;
                ColArgs cn = col_base(p);
                cn.in = Wb; cn.out = Wb; cn.cphi = (float)(q->gamma * q->hz);
                if (launch_col<C, 1, COL_NLSE>(p->q1, cn, st)) return 1;  // :224-228
            }
            if (launch_row(p->q2, Wb, p->T1, p->tw2, N1, N1, st)) return 1;  // :229 of the last step
            ColArgs ci = col_base(p);
            ci.in = Wb; ci.out = E;
            ci.out_scale = (q->amp_mode == OCB_AMP_EDFA) ? (float)sqrt(q->edfa_gain_lin) : gain;
            if (launch_col<C, 1, COL_INV>(p->q1, ci, st)) return 1;  // :232 + gain of :234/:236
            if (q->amp_mode == OCB_AMP_EDFA) {  // noise only; the gain rode in the inverse column pass
                if (launch_amp(E, 1, N, 1.0, inj ? 0.0 : sqrt(q->edfa_noise_var / 2.0),
                               inj ? (const float2*)noise_dev : nullptr, 1, q->seed, (uint64_t)span, st)) return 1;
            }
        } else if (q->amp_mode == OCB_AMP_EDFA) {
            if (launch_amp(E, 1, N, sqrt(q->edfa_gain_lin), inj ? 0.0 : sqrt(q->edfa_noise_var / 2.0),
                           inj ? (const float2*)noise_dev : nullptr, 1, q->seed, (uint64_t)span, st)) return 1;
        }
    }
    return 0;
}
