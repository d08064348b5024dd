// thb_host.cpp -- see thb_host.hpp.  Host arithmetic only; every formula cites the reference
// line it reproduces so that the integer results (hop, win, n_fft, T, n_mel, i_freq_range) are
// exact and the f32 tables (window, mel weights) are the reference's values.
#include "thb_host.hpp"

#include <cfloat>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>

namespace thb {

bool is_pow2(uint64_t x) { return x && !(x & (x - 1)); }

static uint64_t next_pow2(uint64_t x) {  // usize::next_power_of_two (0 -> 1)
    uint64_t p = 1;
    while (p < x) p <<= 1;
    return p;
}

Framing framing_params(const thb_setting &s, uint32_t sr) {
    Framing f;
    // hop = round(win_ms * sr / 1000 / t_overlap) with f64 round-half-away (spectrogram.rs:62-64,91-93)
    const double win_float = s.win_ms * static_cast<double>(sr) / 1000.0;
    const double h = std::round(win_float / static_cast<double>(s.t_overlap));
    f.hop = (h > 0.0 && h == h) ? (h >= 1.8446744073709552e19 ? UINT64_MAX : static_cast<uint64_t>(h)) : 0;
    f.win = f.hop * static_cast<uint64_t>(s.t_overlap);                   // :57-59
    f.n_fft = next_pow2(f.win) * static_cast<uint64_t>(s.f_overlap);      // :95-97
    return f;
}

uint64_t n_frames(uint64_t len, uint64_t win, uint64_t hop) {
    // perform_stft frames the signal reflect-padded by win/2 on both sides with stride hop
    // (stft.rs:50-95); the three pieces together are exactly the windows of that padded signal.
    if (hop == 0) return 0;
    const uint64_t padded = len + 2 * (win / 2);
    return padded >= win ? (padded - win) / hop + 1 : 0;
}

std::vector<float> normalized_hann(uint64_t win, uint64_t n_fft) {
    // cosine_window(0.5, 0.5, 0, 0, win, symmetric = false) / n_fft, all f32 (windows.rs:25,31-38,68-83)
    std::vector<float> w(win);
    const float pi = static_cast<float>(M_PI);
    const float denom = static_cast<float>(win);  // size2 - 1 with size2 = win + 1
    const float nf = static_cast<float>(n_fft);
    for (uint64_t i = 0; i < win; i++) {
        const float x = pi * static_cast<float>(i) / denom;
        const float b_ = 0.5f * cosf(2.0f * x);
        const float c_ = 0.0f * cosf(4.0f * x);
        const float d_ = 0.0f * cosf(6.0f * x);
        const float v = (0.5f - b_) + (c_ - d_);
        w[i] = v / nf;
    }
    return w;
}

// Slaney-style mel <-> Hz in f32 (lib.rs:11-43): constants are f64 literals cast to f32.
static const float kMinLogMel = 15.0f;
static const float kMinLogHz = static_cast<float>(1000.);
static const float kLogStep = static_cast<float>(0.06875177742094912);
static const float kLinearScale = static_cast<float>(200. / 3.);

float mel_to_hz(float mel) {
    if (mel < kMinLogMel) return kLinearScale * mel;
    return kMinLogHz * expf(kLogStep * (mel - kMinLogMel));
}
float mel_from_hz(float hz) {
    if (hz < kMinLogHz) return hz / kLinearScale;
    return kMinLogMel + logf(hz / kMinLogHz) / kLogStep;
}

// calc_mel_fb(sr, n_fft, n_mel, 0, None, true) built band by band in sparse form.
static MelBank build_bank(uint32_t sr, uint64_t n_fft, uint32_t n_mel, bool *all_nonempty) {
    MelBank b;
    b.n_freq = static_cast<uint32_t>(n_fft / 2 + 1);
    b.n_mel = n_mel;
    b.k0.assign(n_mel, 0);
    b.ptr.assign(n_mel + 1, 0);
    const float f_nyquist = static_cast<float>(static_cast<double>(sr) / 2.);
    const uint32_t F = b.n_freq;
    // Array::linspace(0, f_nyquist, F)[i] = 0 + step * i  (lib.rs:64)
    const float step = F > 1 ? (f_nyquist - 0.0f) / static_cast<float>(F - 1) : 0.0f;
    // linspace(from_hz(0), from_hz(fmax), n_mel + 2) mapped through to_hz (lib.rs:65-66)
    std::vector<float> edge(n_mel + 2);
    {
        const float m0 = mel_from_hz(0.0f), m1 = mel_from_hz(f_nyquist);
        const float mstep = (m1 - m0) / static_cast<float>(n_mel + 1);
        for (uint32_t i = 0; i < n_mel + 2; i++) edge[i] = mel_to_hz(m0 + mstep * static_cast<float>(i));
    }
    bool ok = true;
    std::vector<float> wrow;
    for (uint32_t m = 0; m < n_mel; m++) {
        const float e0 = edge[m], e1 = edge[m + 1], e2 = edge[m + 2];
        // first bin with f > e0 (the reference `continue`s over f <= e0, lib.rs:71-72)
        uint64_t k = 0;
        if (step > 0.0f) {
            const double est = static_cast<double>(e0) / static_cast<double>(step);
            k = est > 2.0 ? static_cast<uint64_t>(est) - 2 : 0;
            if (k >= F) k = F - 1;
            while (k > 0 && step * static_cast<float>(k) > e0) k--;  // guard against estimate overshoot
        }
        while (k < F && step * static_cast<float>(k) <= e0) k++;
        wrow.clear();
        const uint64_t kfirst = k;
        for (; k < F; k++) {
            const float f = step * static_cast<float>(k);
            float v;
            if (e0 < f && f < e1) v = (f - e0) / (e1 - e0);              // rising slope  (:73-74)
            else if (f == e1) v = 1.0f;                                   // (:75-76)
            else if (e1 < f && f < e2) v = (e2 - f) / (e2 - e1);          // falling slope (:77-78)
            else break;                                                   // (:79-81)
            wrow.push_back(v);
        }
        // w /= w.sum().max(epsilon) (lib.rs:84-86).  ndarray sums a contiguous row with eight
        // interleaved partial sums over the full row of F entries; zeros do not change them.
        float p[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const uint64_t body = (static_cast<uint64_t>(F) / 8) * 8;
        float acc = 0.0f;
        for (size_t i = 0; i < wrow.size(); i++) {
            const uint64_t kk = kfirst + i;
            if (kk < body) p[kk & 7] += wrow[i];
        }
        acc = acc + (p[0] + p[4]);
        acc = acc + (p[1] + p[5]);
        acc = acc + (p[2] + p[6]);
        acc = acc + (p[3] + p[7]);
        for (size_t i = 0; i < wrow.size(); i++) {
            const uint64_t kk = kfirst + i;
            if (kk >= body) acc = acc + wrow[i];
        }
        float s = acc;
        if (!(s > FLT_EPSILON)) s = FLT_EPSILON;
        bool any = false;
        // trim zero weights at both ends (a slope value can be exactly 0 only by underflow)
        size_t lo = 0, hi = wrow.size();
        for (size_t i = 0; i < wrow.size(); i++) wrow[i] = wrow[i] / s;
        while (lo < hi && wrow[lo] == 0.0f) lo++;
        while (hi > lo && wrow[hi - 1] == 0.0f) hi--;
        b.k0[m] = static_cast<uint32_t>(kfirst + lo);
        for (size_t i = lo; i < hi; i++) {
            b.w.push_back(wrow[i]);
            if (wrow[i] > 0.0f) any = true;
        }
        b.ptr[m + 1] = static_cast<uint32_t>(b.w.size());
        if (!any) ok = false;
    }
    if (all_nonempty) *all_nonempty = ok;
    return b;
}

MelBank mel_bank(uint32_t sr, uint64_t n_fft, uint32_t n_mel) {
    if (n_mel != 0) return build_bank(sr, n_fft, n_mel, nullptr);
    // calc_mel_fb_default (lib.rs:91-103)
    const float v = fmaf(mel_from_hz(static_cast<float>(sr) / 2.0f) /
                             mel_from_hz(static_cast<float>(sr) / static_cast<float>(n_fft)),
                         2.0f, -1.0f);
    uint64_t n = (v > 0.0f && v == v) ? static_cast<uint64_t>(v) : 0;  // `as usize`
    const uint64_t n_freq = n_fft / 2 + 1;
    if (n > n_freq) n = n_freq;
    for (;;) {
        bool ok = false;
        MelBank b = build_bank(sr, n_fft, static_cast<uint32_t>(n), &ok);
        if (ok || n == 0) return b;  // n == 0: `all` over an empty iterator is true in the reference
        n -= 1;
    }
}

MelItems mel_items(const MelBank &b, uint32_t t_multiple, bool with_direct, bool direct_pairs) {
    MelItems it;
    const uint32_t tm = t_multiple < 2 ? 2 : t_multiple;  // the walk takes two steps per float4 of weights
    it.n_mel = b.n_mel;
    const uint32_t F = b.n_freq, M = b.n_mel;
    // ---- per bin: segment id and its two weights (rise of band seg, fall of band seg - 1) ----
    std::vector<int32_t> seg(F, -1);
    std::vector<float> wr(F, 0.0f), wf(F, 0.0f);
    // A bin between the peaks of bands m - 1 and m ("segment" m) carries the falling weight of band m - 1 and the
    // rising weight of band m; a bin inside a single band joins the segment of the bin before it when it can.
    std::vector<int32_t> band_lo(F, -1), band_hi(F, -1);
    std::vector<float> w_lo(F, 0.0f), w_hi(F, 0.0f);
    for (uint32_t m = 0; m < M; m++)
        for (uint32_t i = b.ptr[m]; i < b.ptr[m + 1]; i++) {
            const uint32_t k = b.k0[m] + (i - b.ptr[m]);
            if (k >= F) return it;
            if (band_lo[k] < 0) {
                band_lo[k] = static_cast<int32_t>(m);
                w_lo[k] = b.w[i];
            } else if (band_hi[k] < 0 && band_lo[k] + 1 == static_cast<int32_t>(m)) {
                band_hi[k] = static_cast<int32_t>(m);
                w_hi[k] = b.w[i];
            } else {
                return it;  // not a chain of overlapping triangles: no schedule (the generic kernel handles it)
            }
        }
    for (uint32_t k = 0; k < F; k++) {
        if (band_lo[k] < 0) continue;
        if (band_hi[k] >= 0) {
            seg[k] = band_hi[k];
            wf[k] = w_lo[k];
            wr[k] = w_hi[k];
        } else if (k > 0 && seg[k - 1] == band_lo[k] + 1) {
            seg[k] = band_lo[k] + 1;
            wf[k] = w_lo[k];
        } else {
            seg[k] = band_lo[k];
            wr[k] = w_lo[k];
        }
    }
    // ---- pieces: runs of one segment cut to at most kMelPieceMax bins ----
    struct Piece {
        uint32_t seg, k, len;
    };
    std::vector<Piece> pieces;
    for (uint32_t k = 0; k < F;) {
        if (seg[k] < 0) {
            k++;
            continue;
        }
        uint32_t e = k;
        while (e < F && seg[e] == seg[k]) e++;
        const uint32_t len = e - k, np = (len + kMelPieceMax - 1) / kMelPieceMax;
        const uint32_t base = len / np, rem = len % np;
        uint32_t pos = k;
        for (uint32_t i = 0; i < np; i++) {
            const uint32_t li = base + (i < rem ? 1 : 0);
            pieces.push_back({static_cast<uint32_t>(seg[k]), pos, li});
            pos += li;
        }
        k = e;
    }
    if (pieces.empty()) return it;
    std::vector<uint32_t> order(pieces.size());
    for (uint32_t i = 0; i < order.size(); i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return pieces[x].len > pieces[y].len; });
    it.n_groups = static_cast<uint32_t>((pieces.size() + 31) / 32);
    it.T.resize(it.n_groups);
    it.woff.resize(it.n_groups);
    it.start.assign(static_cast<size_t>(it.n_groups) * 32, 0);
    std::vector<uint32_t> slot_of(pieces.size());
    for (uint32_t g = 0; g < it.n_groups; g++) {
        const uint32_t n = static_cast<uint32_t>(std::min<size_t>(32, order.size() - static_cast<size_t>(g) * 32));
        const uint32_t *ids = &order[static_cast<size_t>(g) * 32];
        uint32_t maxL = 1;
        for (uint32_t i = 0; i < n; i++) maxL = std::max(maxL, pieces[ids[i]].len);
        // Bank assignment.  A float2 load of a half warp is conflict-free when its 16 lanes hit 16 different
        // 8-byte bank pairs, i.e. when their start bins differ mod 16.  A piece shorter than the group's step
        // count may start up to (T - len) bins early on zero weights, which gives it a choice of residues:
        // match pieces to (half warp, residue) slots (Kuhn's augmenting paths), growing T by at most kSlack.
        constexpr uint32_t kSlack = 3;
        int owner[32];      // slot (16 h + r) -> piece index in ids, or -1
        int slot_sel[32];   // piece -> slot
        uint32_t T = maxL;
        bool matched = false;
        for (uint32_t e = 0; e <= kSlack && !matched; e++) {
            T = (maxL + e + tm - 1) / tm * tm;
            std::fill(owner, owner + 32, -1);
            bool seen[32];
            std::function<bool(int)> aug = [&](int i) -> bool {
                const Piece &pc = pieces[ids[i]];
                const uint32_t omax = std::min<uint32_t>(15, T - pc.len);
                for (uint32_t o = 0; o <= omax; o++) {
                    const uint32_t r = (pc.k + 16 * 16 - o) % 16;
                    for (uint32_t h = 0; h < 2; h++) {
                        const uint32_t sl = 16 * h + r;
                        if (seen[sl]) continue;
                        seen[sl] = true;
                        if (owner[sl] < 0 || aug(owner[sl])) {
                            owner[sl] = i;
                            return true;
                        }
                    }
                }
                return false;
            };
            matched = true;
            for (uint32_t i = 0; i < n && matched; i++) {
                std::fill(seen, seen + 32, false);
                matched = aug(static_cast<int>(i));
            }
        }
        if (matched) {
            for (int sl = 0; sl < 32; sl++)
                if (owner[sl] >= 0) slot_sel[owner[sl]] = sl;
        } else {  // give up on conflict freedom for this group: no lead, halves in order
            T = (maxL + tm - 1) / tm * tm;
            for (uint32_t i = 0; i < n; i++) slot_sel[i] = -1;
        }
        // Lane inside the half warp: the lane whose index equals the segment id mod 16 when it is free, so that
        // the gather of consecutive bands reads consecutive bank pairs.
        int lane_of[32];
        bool lane_used[32] = {};
        uint32_t lead[32] = {};
        for (int pass = 0; pass < 2; pass++)
            for (uint32_t i = 0; i < n; i++) {
                const Piece &pc = pieces[ids[i]];
                const int h = matched ? slot_sel[i] / 16 : static_cast<int>(i / 16);
                const int want = 16 * h + static_cast<int>(pc.seg % 16);
                if (pass == 0) {
                    lane_of[i] = -1;
                    if (!lane_used[want]) {
                        lane_of[i] = want;
                        lane_used[want] = true;
                    }
                } else if (lane_of[i] < 0) {
                    int l = 16 * h;
                    while (lane_used[l]) l++;
                    lane_of[i] = l;
                    lane_used[l] = true;
                }
            }
        for (uint32_t i = 0; i < n; i++) {
            const Piece &pc = pieces[ids[i]];
            if (matched) lead[i] = (pc.k + 16 * 16 - static_cast<uint32_t>(slot_sel[i] % 16)) % 16;
        }
        it.T[g] = T;
        it.woff[g] = static_cast<uint32_t>(it.w.size());
        it.w.resize(it.w.size() + static_cast<size_t>(T) * 64, 0.0f);
        for (uint32_t l = 0; l < 32; l++) it.start[static_cast<size_t>(g) * 32 + l] = static_cast<int32_t>(l);  // idle lanes
        for (uint32_t i = 0; i < n; i++) {
            const Piece &pc = pieces[ids[i]];
            const size_t oi = static_cast<size_t>(g) * 32 + static_cast<size_t>(lane_of[i]);
            it.start[oi] = static_cast<int32_t>(pc.k) - static_cast<int32_t>(lead[i]);
            slot_of[ids[i]] = static_cast<uint32_t>(oi);
            for (uint32_t j = 0; j < pc.len; j++) {
                const size_t at = it.w_index(g, lead[i] + j, static_cast<uint32_t>(lane_of[i]));
                it.w[at] = wr[pc.k + j];
                it.w[at + 1] = wf[pc.k + j];
            }
        }
        for (uint32_t l = 0; l < 32; l++) {
            const size_t oi = static_cast<size_t>(g) * 32 + l;
            it.min_start = std::min(it.min_start, it.start[oi]);
            const int64_t reach = static_cast<int64_t>(it.start[oi]) + T - 1;
            if (reach > static_cast<int64_t>(it.max_reach)) it.max_reach = static_cast<uint32_t>(reach);
        }
    }
    // ---- gather lists: band m = rise sums of segment m's pieces, then fall sums of segment m + 1's ----
    const uint32_t n_slots = it.n_groups * 32;
    it.piece_ptr.assign(M + 1, 0);
    for (uint32_t m = 0; m < M; m++) {
        for (uint32_t pi = 0; pi < pieces.size(); pi++)
            if (pieces[pi].seg == m) it.piece_ids.push_back(slot_of[pi]);
        for (uint32_t pi = 0; pi < pieces.size(); pi++)
            if (pieces[pi].seg == m + 1) it.piece_ids.push_back(n_slots + slot_of[pi]);
        it.piece_ptr[m + 1] = static_cast<uint32_t>(it.piece_ids.size());
    }
    // ---- the same lists in the shape the kernels read: per round of 32 bands a uniform count K and K rows of 32
    //      slot ids (u16), short lists padded with a slot that always holds zero ----
    it.zero_slot = 2 * n_slots;
    const uint32_t n_rounds = (M + 31) / 32;
    it.gk.assign(n_rounds, 0);
    it.gbase.assign(n_rounds, 0);
    uint32_t rows = 0;
    for (uint32_t r = 0; r < n_rounds; r++) {
        uint32_t K = 0;
        for (uint32_t m = 32 * r; m < std::min(M, 32 * r + 32); m++) K = std::max(K, it.piece_ptr[m + 1] - it.piece_ptr[m]);
        it.gk[r] = K;
        it.gbase[r] = rows;
        rows += K;
    }
    if (it.zero_slot > 0xffffu) return it;
    it.goff.assign(static_cast<size_t>(rows) * 32, static_cast<uint16_t>(it.zero_slot));
    for (uint32_t m = 0; m < M; m++)
        for (uint32_t i = it.piece_ptr[m]; i < it.piece_ptr[m + 1]; i++)
            it.goff[(static_cast<size_t>(it.gbase[m / 32]) + (i - it.piece_ptr[m])) * 32 + m % 32] = static_cast<uint16_t>(it.piece_ids[i]);
    // rows of four byte offsets (one 16-byte load serves four list entries)
    it.gk4.assign(n_rounds, 0);
    it.gbase4.assign(n_rounds, 0);
    uint32_t rows4 = 0;
    for (uint32_t r = 0; r < n_rounds; r++) {
        it.gk4[r] = (it.gk[r] + 3) / 4;
        it.gbase4[r] = rows4;
        rows4 += it.gk4[r];
    }
    it.goff4.assign(static_cast<size_t>(rows4) * 128, it.zero_slot * 8u);
    for (uint32_t m = 0; m < M; m++)
        for (uint32_t i = it.piece_ptr[m]; i < it.piece_ptr[m + 1]; i++) {
            const uint32_t j = i - it.piece_ptr[m];
            it.goff4[(static_cast<size_t>(it.gbase4[m / 32]) + j / 4) * 128 + (m % 32) * 4 + j % 4] = it.piece_ids[i] * 8u;
        }
    // ---- the band-major schedule and the choice between the two ----
    if (!with_direct) {
        it.valid = true;
        return it;
    }
    // direct_pairs (thb_stft_warp.cu walks two rounds at a time: two independent load -> FMA chains per lane, mel_direct2
    // in thb_stft2048.cuh): both rounds of a pair get the pair's longest band as their step count, and an odd count of
    // rounds is completed with a round of zero weights whose bands do not exist.  The zero steps are shared-memory
    // loads like any other: worth it where the walk is latency bound (n_fft 1024 / 512: -10 %), not in the n_fft 2048
    // frame-pair kernel, whose shared-memory pipe is its busiest unit (+14 % at the 48 kHz default bank).
    const uint32_t d_rounds = direct_pairs ? (n_rounds + 1) & ~1u : n_rounds;
    it.direct_L.assign(d_rounds, 0);
    it.direct_woff.assign(d_rounds, 0);
    it.direct_k0.assign(static_cast<size_t>(d_rounds) * 32, 0);
    uint32_t direct_steps = 0;
    for (uint32_t r = 0; r < d_rounds; r++) {
        const uint32_t m_lo = direct_pairs ? 32 * (r & ~1u) : 32 * r, m_hi = std::min(M, m_lo + (direct_pairs ? 64u : 32u));
        uint32_t L = 0;
        for (uint32_t m = m_lo; m < m_hi; m++) L = std::max(L, b.ptr[m + 1] - b.ptr[m]);
        L = (L + 3) & ~3u;  // four steps per trip
        it.direct_L[r] = L;
        it.direct_woff[r] = static_cast<uint32_t>(it.direct_w.size());
        it.direct_w.resize(it.direct_w.size() + static_cast<size_t>(L) * 32, 0.0f);
        direct_steps += L;
        for (uint32_t m = 32 * r; m < std::min(M, 32 * r + 32); m++) {
            it.direct_k0[m] = static_cast<int32_t>(b.k0[m]);
            for (uint32_t i = b.ptr[m]; i < b.ptr[m + 1]; i++)
                it.direct_w[it.direct_woff[r] + static_cast<size_t>(i - b.ptr[m]) * 32 + (m - 32 * r)] = b.w[i];
            it.direct_reach = std::max<uint32_t>(it.direct_reach, b.k0[m] + L - 1);
        }
        it.direct_reach = std::max<uint32_t>(it.direct_reach, L ? L - 1 : 0);   // lanes without a band read bins 0 .. L - 1
    }
    // Cost of a frame pair in issue slots.  bin-major, from the kernels' SASS: 36 per group + 4.75 per step + per round
    // 12 + 13 per row of four gather entries.  band-major: 10 per round + 4 per step -- the walk is
    // a short dependent chain per lane, so a step costs more than its 3.5 instructions (measured on B200, DESIGN.md
    // section 4: 8.6 per step with one step per trip, ~5 with four, two rounds at a time since).  The band-major schedule
    // is chosen only when it is clearly cheaper: a bank of WIDE bands (mel 128 at 48 kHz: up to 64 bins per band, one lane
    // each) is 13 % slower on it.
    uint32_t steps = 0, rows_of4 = 0;
    for (uint32_t g = 0; g < it.n_groups; g++) steps += it.T[g];
    for (uint32_t r = 0; r < n_rounds; r++) rows_of4 += it.gk4[r];
    const double cost_bin = 36.0 * it.n_groups + 4.75 * steps + 12.0 * n_rounds + 13.0 * rows_of4;
    const double cost_band = 10.0 * n_rounds + (direct_pairs ? 4.0 : 5.0) * direct_steps;
    it.use_direct = cost_band < 0.75 * cost_bin;
    it.valid = true;
    return it;
}

std::vector<uint32_t> MelItems::blob() const {
    // header: {n_groups, n_mel, off_groups, off_start, off_rounds, off_goff, zero_slot, off_w,
    //          off_rounds4, off_goff4, use_direct, off_drounds, off_dk0, off_dw, 0, 0}
    // Only the arrays of the schedule in use are stored (the blob lives in shared memory next to the warps' tiles).
    std::vector<uint32_t> o(16, 0);
    auto align4 = [&]() {
        while (o.size() & 3) o.push_back(0);
    };
    const bool bin_major = !use_direct;
    o[0] = bin_major ? n_groups : 0;
    o[1] = n_mel;
    o[6] = zero_slot;
    o[10] = use_direct ? 1u : 0u;
    o[2] = static_cast<uint32_t>(o.size());  // {T, absolute word offset of the group's weights}
    const size_t grp_at = o.size();
    if (bin_major)
        for (uint32_t g = 0; g < n_groups; g++) {
            o.push_back(T[g]);
            o.push_back(woff[g]);
        }
    align4();
    o[3] = static_cast<uint32_t>(o.size());
    if (bin_major)
        for (int32_t v : start) o.push_back(static_cast<uint32_t>(v));
    align4();
    o[4] = static_cast<uint32_t>(o.size());  // {K, first row} per round of 32 bands
    if (bin_major)
        for (size_t r = 0; r < gk.size(); r++) {
            o.push_back(gk[r]);
            o.push_back(gbase[r]);
        }
    align4();
    o[5] = static_cast<uint32_t>(o.size());
    if (bin_major)
        for (size_t i = 0; i + 1 < goff.size() + 1; i += 2)
            o.push_back(static_cast<uint32_t>(goff[i]) | (static_cast<uint32_t>(i + 1 < goff.size() ? goff[i + 1] : 0) << 16));
    align4();
    o[8] = static_cast<uint32_t>(o.size());  // {rows of four, first row of four} per round
    if (bin_major)
        for (size_t r = 0; r < gk4.size(); r++) {
            o.push_back(gk4[r]);
            o.push_back(gbase4[r]);
        }
    align4();
    o[9] = static_cast<uint32_t>(o.size());
    if (bin_major) o.insert(o.end(), goff4.begin(), goff4.end());
    align4();
    o[11] = static_cast<uint32_t>(o.size());  // {steps, absolute word offset of the round's weights} per round
    const size_t dr_at = o.size();
    if (use_direct)
        for (size_t r = 0; r < direct_L.size(); r++) {
            o.push_back(direct_L[r]);
            o.push_back(direct_woff[r]);
        }
    align4();
    o[12] = static_cast<uint32_t>(o.size());
    if (use_direct)
        for (int32_t v : direct_k0) o.push_back(static_cast<uint32_t>(v));
    align4();
    o[13] = static_cast<uint32_t>(o.size());
    if (use_direct) {
        const uint32_t *dp = reinterpret_cast<const uint32_t *>(direct_w.data());
        o.insert(o.end(), dp, dp + direct_w.size());
        for (size_t r = 0; r < direct_L.size(); r++) o[dr_at + 2 * r + 1] += o[13];
    }
    align4();
    o[7] = static_cast<uint32_t>(o.size());
    if (bin_major) {
        const uint32_t *wp = reinterpret_cast<const uint32_t *>(w.data());
        o.insert(o.end(), wp, wp + w.size());
        for (uint32_t g = 0; g < n_groups; g++) o[grp_at + 2 * g + 1] += o[7];
    }
    align4();
    return o;
}

std::vector<float> MelBank::dense() const {
    std::vector<float> d(static_cast<size_t>(n_freq) * n_mel, 0.0f);
    for (uint32_t m = 0; m < n_mel; m++)
        for (uint32_t i = ptr[m]; i < ptr[m + 1]; i++)
            d[static_cast<size_t>(k0[m] + (i - ptr[m])) * n_mel + m] = w[i];
    return d;
}

void hz_range_to_idx(uint32_t freq_scale, float hz0, float hz1, uint32_t sr, uint64_t n_bins,
                     uint64_t *i0, uint64_t *i1) {
    if (hz0 >= hz1) {  // lib.rs:150-152
        *i0 = 0;
        *i1 = 0;
        return;
    }
    const float half_sr = static_cast<float>(sr) / 2.0f;
    float r0, r1;
    if (freq_scale == THB_FREQ_LINEAR) {  // calc_ratio_to_max_freq (lib.rs:135-141)
        r0 = hz0 / half_sr;
        r1 = hz1 / half_sr;
    } else {
        const float d = mel_from_hz(half_sr);
        r0 = mel_from_hz(hz0) / d;
        r1 = mel_from_hz(hz1) / d;
    }
    float a = floorf(r0 * static_cast<float>(n_bins));
    if (!(a > 0.0f)) a = 0.0f;
    const float c = ceilf(r1 * static_cast<float>(n_bins));
    *i0 = static_cast<uint64_t>(a);
    *i1 = c > 0.0f ? static_cast<uint64_t>(c) : 0;
}

// ---- spectrogram tiles --------------------------------------------------------------------------------------
namespace {
constexpr uint64_t kSpecTile = 512, kSpecGutter = 4;  // SPECTROGRAM_TILE_SIZE / _GUTTER (render_tiles.rs:15-16)
uint64_t sat_mul(uint64_t a, uint64_t b) { return (b && a > UINT64_MAX / b) ? UINT64_MAX : a * b; }
uint64_t sat_sub(uint64_t a, uint64_t b) { return a > b ? a - b : 0; }
double lanczos3(double x) {  // truncated sinc * sinc(x / 3)
    if (!(x >= -3.0 && x < 3.0)) return 0.0;
    auto sinc = [](double t) {
        if (t == 0.0) return 1.0;
        t *= M_PI;
        return std::sin(t) / t;
    };
    return sinc(x) * sinc(x / 3.0);
}
}  // namespace

TileGeometry spectrogram_tile_geometry(uint64_t H, uint64_t W, uint32_t level_x, uint32_t level_y, uint32_t tile_x,
                                       uint32_t tile_y) {
    TileGeometry g;
    const uint64_t sx = level_x < 64 ? (uint64_t(1) << level_x) : UINT64_MAX;  // checked_shl(..).unwrap_or(MAX)
    const uint64_t sy = level_y < 64 ? (uint64_t(1) << level_y) : UINT64_MAX;
    g.lod_width = W / sx + (W % sx != 0);                                       // div_ceil
    g.lod_height = H / sy + (H % sy != 0);
    const uint64_t start_x = sat_mul(tile_x, kSpecTile), start_y = sat_mul(tile_y, kSpecTile);
    const uint64_t core_w = std::min(sat_sub(g.lod_width, start_x), kSpecTile);
    const uint64_t core_h = std::min(sat_sub(g.lod_height, start_y), kSpecTile);
    g.origin_x = sat_sub(start_x, kSpecGutter);
    g.origin_y = sat_sub(start_y, kSpecGutter);
    if (core_w && core_h) {
        g.width = sat_sub(std::min(g.lod_width, start_x + core_w + kSpecGutter), g.origin_x);
        g.height = sat_sub(std::min(g.lod_height, start_y + core_h + kSpecGutter), g.origin_y);
    }
    return g;
}

// fast_image_resize 6.0.0, convolution resize of one axis for U16 pixels (third-party arithmetic, restated from the
// crate's published algorithm -- see oracle/thesia_oracle.c for the statement of what is and is not pinned):
// filter stretched by max(scale, 1); bounds without leading / trailing zero weights; weights normalised to sum 1 in
// f64, then quantised to i32 with the largest precision that keeps the largest weight inside i32.
ResizeAxis resize_axis(uint32_t in_size, double in0, double in1, uint32_t out_size) {
    ResizeAxis a;
    a.n = out_size;
    const double scale = (in1 - in0) / static_cast<double>(out_size);
    const double filter_scale = scale > 1.0 ? scale : 1.0;
    const double radius = 3.0 * filter_scale;
    a.window = static_cast<uint32_t>(std::ceil(radius)) * 2u + 1u;
    const double recip = 1.0 / filter_scale;
    std::vector<double> c(static_cast<size_t>(a.window) * out_size, 0.0);
    a.start.assign(out_size, 0);
    a.size.assign(out_size, 0);
    double max_w = 0.0;
    for (uint32_t o = 0; o < out_size; o++) {
        const double in_center = in0 + (static_cast<double>(o) + 0.5) * scale;
        const double lo = std::max(std::floor(in_center - radius), 0.0);
        const double hi = std::min(std::ceil(in_center + radius), static_cast<double>(in_size));
        const uint32_t x_min = static_cast<uint32_t>(lo), x_max = static_cast<uint32_t>(hi);
        const double center = in_center - 0.5;
        double *row = c.data() + static_cast<size_t>(o) * a.window;
        uint32_t n = 0, b0 = x_min, b1 = x_max;
        double ww = 0.0;
        for (uint32_t x = x_min; x < x_max; x++) {
            const double w = lanczos3((static_cast<double>(x) - center) * recip);
            if (x == b0 && w == 0.0) {
                b0++;
            } else {
                row[n++] = w;
                ww += w;
            }
        }
        for (uint32_t i = n; i-- > 0;) {
            if (b1 <= b0 || row[i] != 0.0) break;
            b1--;
        }
        if (ww != 0.0)
            for (uint32_t i = 0; i < n; i++) row[i] /= ww;
        a.start[o] = b0;
        a.size[o] = b1 - b0;
        for (uint32_t i = 0; i < a.window; i++) max_w = std::max(max_w, row[i]);
    }
    a.precision = 0;
    for (uint32_t p = 0; p < 31; p++) {
        a.precision = p;
        if (std::round(max_w * static_cast<double>(int64_t(1) << (p + 1))) >= 2147483648.0) break;
    }
    const double q = static_cast<double>(int64_t(1) << a.precision);
    a.w_t.assign(static_cast<size_t>(a.window) * out_size, 0);
    std::vector<int32_t> k(a.window);
    for (uint32_t o = 0; o < out_size; o++) {
        for (uint32_t i = 0; i < a.window; i++) k[i] = static_cast<int32_t>(std::round(c[static_cast<size_t>(o) * a.window + i] * q));
        // taps whose weight quantised to zero add nothing to the i64 sum: drop them at both ends of the bound (the
        // result is unchanged bit for bit; at scale 1 on a whole-pixel origin this leaves the single unit tap)
        uint32_t lead = 0, n = a.size[o];
        while (n && k[lead + n - 1] == 0) n--;
        while (n && k[lead] == 0) {
            lead++;
            n--;
        }
        a.start[o] += lead;
        a.size[o] = n;
        for (uint32_t i = 0; i < n; i++) a.w_t[static_cast<size_t>(i) * out_size + o] = k[lead + i];
    }
    return a;
}

std::vector<float> twiddle_table(uint64_t n_fft) {
    std::vector<float> t(2 * n_fft);
    for (uint64_t i = 0; i < n_fft; i++) {
        const double ang = -2.0 * M_PI * static_cast<double>(i) / static_cast<double>(n_fft);
        t[2 * i] = static_cast<float>(std::cos(ang));
        t[2 * i + 1] = static_cast<float>(std::sin(ang));
    }
    return t;
}

}  // namespace thb
