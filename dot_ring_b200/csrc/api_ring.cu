// C ABI, part 3: ring set-up (key ingestion, fixed columns, ring root) and batched Ring VRF proving.
#include <cstdio>
#include <cstdlib>

#include "api_internal.cuh"
#include "ring.cuh"

namespace dr {

static Fr fr_from_param(const uint8_t* b) {
    Fr r;
    fr_from_le_bytes_raw(r, b);
    if (!r.is_canonical_raw()) throw Error(DR_EINVAL, "field element is not canonical");
    return r.to_mont();
}
static TEAffine te_from_param(const uint8_t* xy) {
    TEAffine p{fr_from_param(xy), fr_from_param(xy + 32)};
    if (!te_on_curve(p)) throw Error(DR_EINVAL, "auxiliary point is not on the curve");
    return p;
}

// Per-pass device scratch.  Buffers with disjoint lifetimes share storage: the 16N-element low-degree extensions are dead once
// the constraints are evaluated, so the combined quotient coefficients (cagg), the aggregated opening polynomial (aggopen) and
// the linearisation polynomial (lin), all born later, live inside that block: 27 N + 1 field elements per proof instead of 35 N.
struct ProveScratch {
    size_t cap = 0;
    uint32_t N = 0;
    DevBuf<ProofState> st;
    DevBuf<ProveInput> in;
    DevBuf<Fr> wit_coef, lde, agg, quot, ntt_tmp;
    struct View {
        Fr* p = nullptr;
    } cagg, aggopen, lin;
    DevBuf<G1Affine> res;
    DevBuf<uint8_t> out, zraw, blob;
    DevBuf<uint32_t> status;
    static size_t elements_per_proof(uint32_t N) { return (size_t)27 * N + 1; }
    void ensure(size_t n, uint32_t N_) {
        if (n <= cap && N_ == N) return;
        cap = n;
        N = N_;
        size_t q = 3 * (size_t)N + 1;
        st.alloc(n);
        in.alloc(n);
        wit_coef.alloc(n * 4 * N);
        lde.alloc(n * 16 * N);
        cagg.p = lde.p;                               // n x 4N
        aggopen.p = lde.p + n * 4 * (size_t)N;        // n x (3N + 1)
        lin.p = lde.p + n * (7 * (size_t)N + 1);      // n x N
        agg.alloc(n * 4 * N);
        quot.alloc(n * q);
        res.alloc(n * 4);
        out.alloc(n * 784);
        zraw.alloc(n * 12 * 32);
        status.alloc(n);
    }
};

static ProveScratch& scratch_for(Ctx* ctx) {
    if (!ctx->prove_scratch) ctx->prove_scratch = std::shared_ptr<void>(new ProveScratch(), [](void* p) { delete (ProveScratch*)p; });
    return *(ProveScratch*)ctx->prove_scratch.get();
}

}  // namespace dr

using namespace dr;

#define DR_API_BEGIN try {
#define DR_API_END                            \
    }                                         \
    catch (const Error& e) {                  \
        return set_error(e.code, e.what());   \
    }                                         \
    catch (const std::exception& e) {         \
        return set_error(DR_ECUDA, e.what()); \
    }                                         \
    return DR_OK;

extern "C" {

int dr_ring_create(dr_ctx* c, dr_srs* s, const dr_ring_params* prm, const uint8_t* keys32, size_t n_keys, dr_ring** out) {
    DR_API_BEGIN
    Ctx* ctx = (Ctx*)c;
    Srs* srs = (Srs*)s;
    if (!ctx || !srs || !prm || !out || (n_keys && !keys32)) throw Error(DR_EINVAL, "bad argument");
    ctx->activate();
    const uint32_t N = prm->domain_size;
    // the reference stops at 4096 (params.py:172-173); larger domains (to 2^16) run the same pipeline over the two-pass NTT
    if (N < 512 || N > 65536 || (N & (N - 1))) throw Error(DR_EINVAL, "domain_size must be a power of two in [512, 65536]");
    if (prm->padding_rows != 4) throw Error(DR_EINVAL, "padding_rows must be 4 to match the 3 hidden rows");
    if (prm->max_ring_size + SCALAR_BITS + 4 > N) throw Error(DR_EINVAL, "max_ring_size exceeds supported size for this domain");
    if (n_keys > prm->max_ring_size) throw Error(DR_EINVAL, "ring size exceeds max supported size");
    if (3 * (size_t)N + 1 > srs->n) throw Error(DR_EINVAL, "SRS too small for this domain (needs 3N+1 G1 points)");
    if (prm->suite_id_len > 32 || prm->h2c_dst_len > 64) throw Error(DR_EINVAL, "suite id / DST too long");

    auto ring = std::make_unique<Ring>();
    ring->ctx = ctx;
    ring->srs = srs;
    RingDev& d = ring->dev;
    d.N = N;
    d.logN = 0;
    while ((1u << d.logN) < N) d.logN++;
    d.max_ring = prm->max_ring_size;
    d.last = N - 4;
    d.omega = fr_from_param(prm->omega);
    Fr w4 = fr_from_param(prm->radix_omega);
    // sanity: w4^4 == omega, omega^N == 1, omega^(N/2) != 1
    if (w4.sqr().sqr() != d.omega) throw Error(DR_EINVAL, "radix_omega^4 != omega");
    {
        Fr t = d.omega;
        for (uint32_t i = 0; i + 1 < d.logN; i++) t = t.sqr();
        if (t == Fr::one() || t.sqr() != Fr::one()) throw Error(DR_EINVAL, "omega is not a primitive N-th root of unity");
    }
    d.seed = te_from_param(prm->seed);
    d.blinding_base = te_from_param(prm->blinding_base);
    d.generator = te_from_param(prm->generator);
    ring->g_table = ctx->fixed_table(d.generator);
    ring->b_table = ctx->fixed_table(d.blinding_base);
    d.g_tab = ring->g_table->tab.p;
    d.b_tab = ring->b_table->tab.p;
    ring->padding = te_from_param(prm->padding_point);
    d.suite_id_len = prm->suite_id_len;
    memcpy(d.suite_id, prm->suite_id, 32);
    d.dst_len = prm->h2c_dst_len;
    memcpy(d.dst, prm->h2c_dst, 64);
    if (prm->hash_id > 1) throw Error(DR_EINVAL, "unknown suite hash");
    d.hash_kind = prm->hash_id;
    d.n_inv = Fr::from_u32(N).inv();
    d.quarter = Fr::from_u32(4).inv();

    // ---- domain tables (host, portable arithmetic; uploaded once) ----
    const NttPlan& plan = ctx->plan(N, d.omega);
    d.tw_fwd = plan.tw_fwd.p;
    d.tw_inv = plan.tw_inv.p;
    std::vector<Fr> hw4(4 * N), hw4i(4 * N);
    {
        Fr a = Fr::one(), b = Fr::one(), w4i = w4.inv();
        for (uint32_t k = 0; k < 4 * N; k++) {
            hw4[k] = a;
            hw4i[k] = b;
            a = a * w4;
            b = b * w4i;
        }
        if (a != Fr::one()) throw Error(DR_EINVAL, "radix_omega is not a 4N-th root of unity");
    }
    ring->w4.alloc(4 * N);
    ring->w4inv.alloc(4 * N);
    h2d(ctx->stream, ring->w4.p, hw4.data(), 4 * N * sizeof(Fr));
    h2d(ctx->stream, ring->w4inv.p, hw4i.data(), 4 * N * sizeof(Fr));
    d.w4 = ring->w4.p;
    d.w4inv = ring->w4inv.p;
    d.w_last = hw4[4 * (N - 4)];
    {
        // (X - w^(N-1))(X - w^(N-2))(X - w^(N-3))
        Fr r1 = hw4[4 * (N - 1)], r2 = hw4[4 * (N - 2)], r3 = hw4[4 * (N - 3)];
        d.tail[3] = Fr::one();
        d.tail[2] = (r1 + r2 + r3).neg();
        d.tail[1] = r1 * r2 + r1 * r3 + r2 * r3;
        d.tail[0] = (r1 * r2 * r3).neg();
    }

    // ---- public vector PK || padding || 2^i B || zeros (members.py:36-53) ----
    ring->nm.alloc(N);
    {
        std::vector<TEAffine> host(N);
        for (uint32_t i = 0; i < d.max_ring; i++) host[i] = ring->padding;
        TEExt cur = TEExt::from_affine(d.blinding_base);
        for (uint32_t i = d.max_ring; i < N - 4; i++) {
            host[i] = te_to_affine(cur);
            cur = te_dbl(cur);
        }
        for (uint32_t i = N - 4; i < N; i++) host[i] = TEAffine{Fr::zero(), Fr::zero()};
        h2d(ctx->stream, ring->nm.p, host.data(), N * sizeof(TEAffine));
        if (n_keys) {
            DevBuf<uint8_t> kraw(n_keys * 32);
            h2d(ctx->stream, kraw.p, keys32, n_keys * 32);
            launch(ctx->stream, Dim3((uint32_t)((n_keys + 63) / 64)), 64, 0, KeyDecodeBody(), (const uint8_t*)kraw.p, (uint32_t)n_keys, ring->padding, ring->nm.p);
            stream_sync(ctx->stream);
        }
    }
    d.nm = ring->nm.p;

    // ---- fixed columns: evaluations -> coefficients -> commitments ----
    ring->fixed_coef.alloc(3 * N);
    launch(ctx->stream, Dim3((N + 127) / 128), 128, 0, FixedColumnsBody(), (const TEAffine*)ring->nm.p, N, d.max_ring, ring->fixed_coef.p);
    const uint32_t nthr = ntt_threads(N);
    DevBuf<Fr> ntt_tmp;
    ntt_device(ctx, plan, ring->fixed_coef.p, ring->fixed_coef.p, 3, true, ntt_tmp);
    d.fixed_coef = ring->fixed_coef.p;
    DevBuf<G1Affine> cm(3);
    commit_device(ctx, srs, ring->fixed_coef.p, N, N, 3, cm.p);
    DevBuf<uint8_t> enc96(288), enc48(144);
    launch(ctx->stream, Dim3(1), 64, 0, G1EncodeBody(), (const G1Affine*)cm.p, 3u, enc96.p, enc48.p);
    d2h(ctx->stream, ring->commitments, cm.p, 3 * sizeof(G1Affine));
    d2h(ctx->stream, ring->commit_be96, enc96.p, 288);
    d2h(ctx->stream, ring->root144, enc48.p, 144);

    // ---- 4x LDE of px, py, s, L_0, L_{N-4} and (x - w^(N-4)) on the 4N domain ----
    ring->fixed_lde.alloc(6 * 4 * (size_t)N);
    {
        DevBuf<Fr> coef5(5 * (size_t)N);
        d2d(ctx->stream, coef5.p, ring->fixed_coef.p, 3 * N * sizeof(Fr));
        std::vector<Fr> lag(2 * (size_t)N);
        Fr cur0 = d.n_inv, curl = d.n_inv;
        Fr inv_xl = hw4i[4 * (N - 4)];  // w^-(N-4)
        for (uint32_t k = 0; k < N; k++) {
            lag[k] = cur0;  // L_0 = (1/N) sum X^k
            lag[N + k] = curl;
            curl = curl * inv_xl;
        }
        h2d(ctx->stream, coef5.p + 3 * (size_t)N, lag.data(), 2 * N * sizeof(Fr));
        if (N <= 4096 && !ctx->generic_ntt_path) {
            launch(ctx->stream, Dim3(20), nthr, ntt_smem_bytes(N), PlainLdeBody(), N, d.logN, (const Fr*)plan.tw_fwd.p, (const Fr*)ring->w4.p, (const Fr*)coef5.p,
                   ring->fixed_lde.p);
        } else {
            launch(ctx->stream, Dim3((N + 127) / 128, 4, 5), 128, 0, CosetTwistBody(), N, (const Fr*)ring->w4.p, (const Fr*)coef5.p, ring->fixed_lde.p);
            ntt_device(ctx, plan, ring->fixed_lde.p, ring->fixed_lde.p, 20, false, ntt_tmp);
        }
        launch(ctx->stream, Dim3((4 * N + 127) / 128), 128, 0, NotLastBody(), N, (const Fr*)ring->w4.p, d.w_last, ring->fixed_lde.p + 5 * 4 * (size_t)N);
        stream_sync(ctx->stream);
    }
    d.fixed_lde = ring->fixed_lde.p;

    // ---- verifier-key transcript prefix (root.py:54-71, phases.py:72-74) ----
    {
        Shake128 tr;
        tr.init();
        tr.absorb(d.suite_id, d.suite_id_len);
        tr.absorb_be32(d.suite_id_len);
        shake_absorb_label(tr, "vk", 2);
        tr.absorb(srs->g1_0_be96, 96);
        tr.absorb(srs->g2_be192, 384);
        tr.absorb(ring->commit_be96, 288);
        tr.absorb_be32(96 + 384 + 288);
        ring->prefix.alloc(1);
        h2d(ctx->stream, ring->prefix.p, &tr, sizeof(Shake128));
        stream_sync(ctx->stream);
    }
    ring->vk = make_verifier_key(ctx, N, d.omega, d.seed, d.suite_id, d.suite_id_len, srs->g1_0_be96, srs->g2_be192, ring->commit_be96, ring->vk_lines);
    ring->suite.generator = d.generator;
    ring->suite.blinding_base = d.blinding_base;
    ring->suite.g_tab = d.g_tab;
    ring->suite.b_tab = d.b_tab;
    ring->suite.suite_id_len = d.suite_id_len;
    memcpy(ring->suite.suite_id, d.suite_id, 32);
    ring->suite.dst_len = d.dst_len;
    memcpy(ring->suite.dst, d.dst, 64);
    ring->suite.hash_kind = d.hash_kind;
    *out = (dr_ring*)ring.release();
    DR_API_END
}

void dr_ring_destroy(dr_ring* r) { delete (Ring*)r; }

int dr_ring_root(dr_ring* r, uint8_t root144[144]) {
    DR_API_BEGIN
    if (!r || !root144) throw Error(DR_EINVAL, "bad argument");
    memcpy(root144, ((Ring*)r)->root144, 144);
    DR_API_END
}

int dr_ring_fixed_commitments(dr_ring* r, uint8_t out288[288]) {
    DR_API_BEGIN
    if (!r || !out288) throw Error(DR_EINVAL, "bad argument");
    memcpy(out288, ((Ring*)r)->commit_be96, 288);
    DR_API_END
}

int dr_ring_points(dr_ring* r, uint8_t* out_xy64, size_t n_points) {
    DR_API_BEGIN
    Ring* ring = (Ring*)r;
    if (!ring || !out_xy64 || n_points > ring->dev.N) throw Error(DR_EINVAL, "bad argument");
    ring->ctx->activate();
    std::vector<TEAffine> host(n_points);
    d2h(ring->ctx->stream, host.data(), ring->nm.p, n_points * sizeof(TEAffine));
    stream_sync(ring->ctx->stream);
    for (size_t i = 0; i < n_points; i++) {
        fr_to_le_bytes_raw(out_xy64 + 64 * i, host[i].x.from_mont());
        fr_to_le_bytes_raw(out_xy64 + 64 * i + 32, host[i].y.from_mont());
    }
    DR_API_END
}

int dr_ring_prove_batch(dr_ctx* c, dr_ring* r, size_t n, const uint8_t* blob, const uint32_t* alpha_off, const uint32_t* alpha_len, const uint32_t* ad_off,
                        const uint32_t* ad_len, const uint8_t* secret_keys32, const uint32_t* producer_index, const uint8_t* zk_rows, uint8_t* proofs784,
                        uint32_t* status) {
    DR_API_BEGIN
    Ctx* ctx = (Ctx*)c;
    Ring* ring = (Ring*)r;
    if (!ctx || !ring || (n && (!alpha_off || !alpha_len || !ad_off || !ad_len || !secret_keys32 || !producer_index || !proofs784 || !status)))
        throw Error(DR_EINVAL, "bad argument");
    ctx->activate();
    if (!n) return DR_OK;
    const RingDev& rg = ring->dev;
    const uint32_t N = rg.N;
    const uint32_t qlen = 3 * N + 1;
    size_t blob_len = 0;
    for (size_t i = 0; i < n; i++) {
        size_t e1 = (size_t)alpha_off[i] + alpha_len[i], e2 = (size_t)ad_off[i] + ad_len[i];
        if (e1 > blob_len) blob_len = e1;
        if (e2 > blob_len) blob_len = e2;
    }
    if (blob_len && !blob) throw Error(DR_EINVAL, "null blob");
    // (the input blob goes into a buffer that lives with the scratch: no allocation on the steady-state path)

    // proofs per pass: the one-thread-per-proof kernels (Pedersen part, witness chain, transcripts) are latency-bound, so
    // a pass should be as wide as the scratch (35 N field elements + state per proof) allows
    size_t chunk_cap = ctx->prove_chunk ? ctx->prove_chunk : 8192;
#if !defined(DR_HOST_EMULATION)
    // The free-memory query (like any allocation) can stall for tens of ms inside the driver, so it is only made when the scratch
    // has to grow: a call that fits the scratch of an earlier one reuses its pass width.
    {
        ProveScratch& cur = scratch_for(ctx);
        const size_t want = n < chunk_cap ? n : chunk_cap;
        if (!ctx->prove_chunk && !(cur.N == N && want <= cur.cap)) {
            size_t free_b = 0, total_b = 0;
            DR_CUDA(cudaMemGetInfo(&free_b, &total_b));
            free_b += dev_cache().cached;  // recycled blocks are available to this call
            size_t per_proof = (ProveScratch::elements_per_proof(N) + (N > 4096 ? 16 * (size_t)N : 0)) * sizeof(Fr) + sizeof(ProofState) + 4096;
            size_t have = free_b + (cur.N == N ? cur.cap * per_proof : 0);
            // keep 1 GB (or half of what is left, if less) for the other buffers of this and later calls
            const size_t keep = have / 2 < ((size_t)1 << 30) ? have / 2 : ((size_t)1 << 30);
            size_t fit = (have - keep) / per_proof;
            if (getenv("DOT_RING_B200_DEBUG"))
                fprintf(stderr, "[dot_ring_b200] prove: free %.2f GB (+%.2f GB held), %.2f MB per proof -> up to %zu proofs per pass\n", free_b / 1e9,
                        (double)(have - free_b) / 1e9, per_proof / 1e6, fit);
            if (fit < chunk_cap) chunk_cap = fit < 64 ? 64 : fit;
        }
    }
#endif
    // equal passes: n proofs in ceil(n / cap) passes of the same width (a short last pass would pay the latency-bound kernels again
    // for little work)
    const size_t passes = (n + chunk_cap - 1) / chunk_cap;
    const size_t pass_width = (n + passes - 1) / passes;
    ProveScratch& sc = scratch_for(ctx);
    sc.ensure(pass_width, N);
    sc.blob.ensure(blob_len ? blob_len : 1);
    h2d(ctx->stream, sc.blob.p, blob, blob_len);
    PhaseTimer& pt = ctx->phases;
    pt.reset();
    pt.active = true;
    struct Deactivate {
        PhaseTimer& p;
        ~Deactivate() { p.active = false; }
    } deactivate{pt};
    const uint32_t nthr = ntt_threads(N);
    const size_t ntt_smem = ntt_smem_bytes(N);
    const bool large = N > 4096 || ctx->generic_ntt_path;  // two-pass transforms + element-wise twists instead of the fused single-CTA kernels
    const NttPlan* big_plan = large ? &ctx->plan(N, rg.omega) : nullptr;
    std::vector<ProveInput> hin;

    for (size_t base = 0; base < n; base += pass_width) {
        const uint32_t m = (uint32_t)((n - base < pass_width) ? n - base : pass_width);
        pt.mark(ctx, 7);
        hin.assign(m, ProveInput{});
        for (uint32_t i = 0; i < m; i++) {
            ProveInput& pi = hin[i];
            pi.alpha_off = alpha_off[base + i];
            pi.alpha_len = alpha_len[base + i];
            pi.ad_off = ad_off[base + i];
            pi.ad_len = ad_len[base + i];
            pi.k = producer_index[base + i];
            memcpy(pi.sk, secret_keys32 + 32 * (base + i), 32);
        }
        h2d(ctx->stream, sc.in.p, hin.data(), m * sizeof(ProveInput));
        const uint32_t tb = 64, pb = (m + tb - 1) / tb;

        pt.mark(ctx, 0);
        // zk rows (Montgomery) into the state, then Pedersen + witness
        if (zk_rows) {
            h2d(ctx->stream, sc.zraw.p, zk_rows + (size_t)base * 12 * 32, (size_t)m * 12 * 32);
            launch(ctx->stream, Dim3((m * 12 + 127) / 128), 128, 0, ZkRowsBody(), (const uint8_t*)sc.zraw.p, sc.st.p, m);
        } else {
            launch(ctx->stream, Dim3((m * 12 + 127) / 128), 128, 0, ZkRowsBody(), (const uint8_t*)nullptr, sc.st.p, m);
        }
        // Pedersen part: the half the ring proof needs (blinding factor, blinded key) on the main stream, eight lanes per proof; the
        // rest (nonces, R, Ok, responses) on the side stream, joined before the proofs are assembled
        {
            const uint32_t per_block = tb / COOP_LANES;  // eight lanes per proof
            launch(ctx->stream, Dim3((m + per_block - 1) / per_block), tb, pedersen_start_smem(tb), PedersenStartBody(), rg, (const ProveInput*)sc.in.p, (const uint8_t*)sc.blob.p,
                   sc.st.p, m);
        }
        ctx->fork_side();
        launch(ctx->side, Dim3((m + 31) / 32), 32, 0, PedersenFinishBody(), rg, (const ProveInput*)sc.in.p, sc.st.p, m);
        {
            const uint32_t wt = 4 * WIT_LANES;  // four proofs per block, a warp each
            launch(ctx->stream, Dim3((m + 3) / 4), wt, witness_coop_smem(wt), WitnessBody(), rg, sc.st.p, m, (const Shake128*)ring->prefix.p);
        }
        pt.mark(ctx, 1);
        // The coefficient form of the witness columns is first needed by the low-degree extension, the commitments by the transcript:
        // on the fused path the interpolation runs on a second side stream next to the (sparse) commitments.
        const bool overlap_intt = !large && !ctx->dense_witness_commit;
        if (!large) {
            if (overlap_intt) ctx->fork_side2();
            launch(overlap_intt ? ctx->side2 : ctx->stream, Dim3(4, m), nthr, ntt_smem, WitnessInttBody(), rg, (const ProofState*)sc.st.p, sc.wit_coef.p);
        } else {
            launch(ctx->stream, Dim3((N + 127) / 128, 4, m), 128, 0, WitnessEvalBody(), rg, (const ProofState*)sc.st.p, sc.wit_coef.p);
            ntt_device(ctx, *big_plan, sc.wit_coef.p, sc.wit_coef.p, 4 * (size_t)m, true, sc.ntt_tmp);
        }
        pt.mark(ctx, 2);
        if (ctx->dense_witness_commit || large) {
            commit_device(ctx, ring->srs, sc.wit_coef.p, N, N, 4 * m, sc.res.p);
        } else {
            if (!ring->lag) ring->lag = &ring->srs->lagrange_table(N, rg.logN, rg.omega, rg.tw_inv, rg.n_inv);
            ctx->partials.ensure((size_t)4 * m);
            launch_lb<64, 8>(ctx->stream, Dim3(4, m), 64, witness_commit_smem(ring->lag->geom.W, 64), WitnessCommitBody(), (const G1Affine*)ring->lag->table.p, ring->lag->geom, rg, (const ProofState*)sc.st.p,
                   ctx->partials.p);
            launch_commit_finish(ctx->stream, (const G1*)ctx->partials.p, 1u, 4 * m, sc.res.p);
        }
        // column order (b, accx, accy, accip) -> payload slots (0, 2, 3, 1)
        launch(ctx->stream, Dim3((4 * m + 127) / 128), 128, 0, StoreCommitBody(), (const G1Affine*)sc.res.p, 4u, 0x01030200u, sc.st.p, m);
        pt.mark(ctx, 5);
        launch(ctx->stream, Dim3(pb), tb, 0, Transcript1Body(), sc.st.p, m);
        if (overlap_intt) ctx->join_side2();
        pt.mark(ctx, 3);
        if (!large) {
            launch(ctx->stream, Dim3(16, m), nthr, ntt_smem, WitnessLdeBody(), rg, (const ProofState*)sc.st.p, (const Fr*)sc.wit_coef.p, sc.lde.p);
        } else {
            launch(ctx->stream, Dim3((N + 127) / 128, 4, 4 * m), 128, 0, CosetTwistBody(), N, rg.w4, (const Fr*)sc.wit_coef.p, sc.lde.p);
            ntt_device(ctx, *big_plan, sc.lde.p, sc.lde.p, 16 * (size_t)m, false, sc.ntt_tmp);
        }
        launch(ctx->stream, Dim3((4 * N + 127) / 128, m), 128, 0, ConstraintBody(), rg, (const ProofState*)sc.st.p, (const Fr*)sc.lde.p, sc.agg.p);
        if (!large) {
            launch(ctx->stream, Dim3(4, m), nthr, ntt_smem, QuotientInttBody(), rg, sc.agg.p);
        } else {
            ntt_device(ctx, *big_plan, sc.agg.p, sc.agg.p, 4 * (size_t)m, true, sc.ntt_tmp);
            launch(ctx->stream, Dim3((N + 127) / 128, 4, m), 128, 0, CosetUntwistBody(), N, rg.w4inv, sc.agg.p);
        }
        launch(ctx->stream, Dim3((N + 127) / 128, m), 128, 0, QuotientCombineBody(), rg, (const Fr*)sc.agg.p, sc.cagg.p);
        launch(ctx->stream, Dim3((qlen + 127) / 128, m), 128, 0, QuotientFoldBody(), rg, (const Fr*)sc.cagg.p, sc.quot.p, qlen);
        pt.mark(ctx, 2);
        commit_device(ctx, ring->srs, sc.quot.p, qlen, qlen, m, sc.res.p);
        launch(ctx->stream, Dim3((m + 127) / 128), 128, 0, StoreCommitBody(), (const G1Affine*)sc.res.p, 1u, 0x04u, sc.st.p, m);
        pt.mark(ctx, 5);
        launch(ctx->stream, Dim3(pb), tb, 0, Transcript2Body(), sc.st.p, m);
        pt.mark(ctx, 4);
        launch(ctx->stream, Dim3(7, m), 128, 128 * sizeof(Fr), EvalBody(), rg, sc.st.p, (const Fr*)sc.wit_coef.p);
        launch(ctx->stream, Dim3((N + 127) / 128, m), 128, 0, LinPolyBody(), rg, (const ProofState*)sc.st.p, (const Fr*)sc.wit_coef.p, sc.lin.p);
        launch(ctx->stream, Dim3(1, m), 128, 128 * sizeof(Fr), LinEvalBody(), rg, sc.st.p, (const Fr*)sc.lin.p);
        pt.mark(ctx, 5);
        launch(ctx->stream, Dim3(pb), tb, 0, Transcript3Body(), sc.st.p, m);
        pt.mark(ctx, 4);
        launch(ctx->stream, Dim3((qlen + 127) / 128, m), 128, 0, AggOpenBody(), rg, (const ProofState*)sc.st.p, (const Fr*)sc.wit_coef.p, (const Fr*)sc.quot.p, qlen, sc.aggopen.p);
        launch(ctx->stream, Dim3(2, m), 256, synthetic_div_smem(256), OpenQuotientsBody(), rg, (const ProofState*)sc.st.p, sc.aggopen.p, qlen, sc.lin.p);
        pt.mark(ctx, 2);
        commit_device_pair(ctx, ring->srs, sc.aggopen.p, qlen, 3 * N, sc.lin.p, N, N - 1, m, sc.res.p);
        launch(ctx->stream, Dim3((m + 127) / 128), 128, 0, StoreCommitBody(), (const G1Affine*)sc.res.p, 1u, 0x05u, sc.st.p, m);
        launch(ctx->stream, Dim3((m + 127) / 128), 128, 0, StoreCommitBody(), (const G1Affine*)(sc.res.p + m), 1u, 0x06u, sc.st.p, m);
        pt.mark(ctx, 5);
        ctx->join_side();
        launch(ctx->stream, Dim3(pb), tb, 0, FinalizeBody(), (const ProofState*)sc.st.p, m, sc.out.p, sc.status.p);
        pt.mark(ctx, 8);
        dev_zero(ctx->stream, sc.in.p, m * sizeof(ProveInput));  // secret keys do not outlive the pass on the device
        d2h(ctx->stream, proofs784 + 784 * base, sc.out.p, (size_t)m * 784);
        d2h(ctx->stream, status + base, sc.status.p, (size_t)m * sizeof(uint32_t));
        pt.mark(ctx, -1);
        stream_sync(ctx->stream);
        pt.collect(ctx);
    }
    {  // ... nor in the host staging buffer (volatile: the stores must not be optimised away)
        volatile uint8_t* w = (volatile uint8_t*)hin.data();
        for (size_t i = 0; i < hin.size() * sizeof(ProveInput); i++) w[i] = 0;
    }
    DR_API_END
}

int dr_ring_prove_phase_ms(dr_ctx* c, float out[6]) {
    DR_API_BEGIN
    Ctx* ctx = (Ctx*)c;
    if (!ctx || !out) throw Error(DR_EINVAL, "bad argument");
    for (int i = 0; i < 6; i++) out[i] = ctx->phases.total[i];
    if (getenv("DOT_RING_B200_DEBUG"))
        fprintf(stderr, "[dot_ring_b200] prove: copy-in / set-up %.3f ms, copy-out %.3f ms, commit kernel %.3f ms, side stream fork -> done %.3f ms\n", ctx->phases.total[7],
                ctx->phases.total[8], ctx->phases.total[6], ctx->side_span_ms());
    DR_API_END
}

int dr_ring_prove_commit_kernel_ms(dr_ctx* c, float* ms, uint32_t* launches) {
    DR_API_BEGIN
    Ctx* ctx = (Ctx*)c;
    if (!ctx || !ms || !launches) throw Error(DR_EINVAL, "bad argument");
    *ms = ctx->phases.total[6];
    *launches = ctx->phases.kernel_launches;
    DR_API_END
}

uint32_t dr_ring_witness_table_bits(const dr_ring* r) {
    const Ring* ring = (const Ring*)r;
    return ring && ring->lag ? ring->lag->geom.c : 0;
}

int dr_ctx_set_generic_ntt_path(dr_ctx* c, int enabled) {
    DR_API_BEGIN
    Ctx* ctx = (Ctx*)c;
    if (!ctx) throw Error(DR_EINVAL, "bad argument");
    ctx->generic_ntt_path = enabled != 0;
    DR_API_END
}

int dr_ctx_set_dense_witness_commit(dr_ctx* c, int enabled) {
    DR_API_BEGIN
    Ctx* ctx = (Ctx*)c;
    if (!ctx) throw Error(DR_EINVAL, "bad argument");
    ctx->dense_witness_commit = enabled != 0;
    DR_API_END
}

int dr_ctx_set_prove_chunk(dr_ctx* c, size_t chunk) {
    DR_API_BEGIN
    Ctx* ctx = (Ctx*)c;
    if (!ctx) throw Error(DR_EINVAL, "bad argument");
    ctx->prove_chunk = chunk;
    DR_API_END
}

}  // extern "C"
