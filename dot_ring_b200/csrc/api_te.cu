// C ABI, part 4: batched Bandersnatch point operations (decode + subgroup check, scalar multiplication).
#include "api_internal.cuh"
#include "ring.cuh"

namespace dr {

struct TeDecodeBody {
    DR_HD void operator()(const BlockCtx& ctx, const uint8_t* in, uint32_t n, int checked, uint8_t* out_xy, uint8_t* ok) const {
        DR_THREAD_LOOP(t, ctx) {
            uint32_t i = ctx.bx * ctx.nthreads + t;
            if (i < n) {
                TEAffine p;
                bool good = checked ? te_decode_checked(p, in + 32 * (size_t)i) : te_decode(p, in + 32 * (size_t)i);
                if (!good) p = TEAffine::identity();
                fr_to_le_bytes_raw(out_xy + 64 * (size_t)i, p.x.from_mont());
                fr_to_le_bytes_raw(out_xy + 64 * (size_t)i + 32, p.y.from_mont());
                ok[i] = good ? 1 : 0;
            }
        }
    }
};

// out[i] = (k_i mod n) * P_(i or 0); ok[i] = 0 when the point does not decode
struct TeMulBody {
    DR_HD void operator()(const BlockCtx& ctx, const uint8_t* pts, uint32_t n_pts, const uint8_t* ks, uint32_t n, uint8_t* out, uint8_t* ok) const {
        DR_THREAD_LOOP(t, ctx) {
            uint32_t i = ctx.bx * ctx.nthreads + t;
            if (i < n) {
                TEAffine p;
                bool good = te_decode(p, pts + 32 * (size_t)(n_pts == 1 ? 0 : i));
                if (good) {
                    Fn k = fp_from_le_bytes_mod<Fn>(ks + 32 * (size_t)i, 32);
                    te_encode(out + 32 * (size_t)i, te_mul_fn(p, k));
                } else {
                    for (int b = 0; b < 32; b++) out[32 * (size_t)i + b] = 0;
                }
                ok[i] = good ? 1 : 0;
            }
        }
    }
};

}  // namespace dr

using namespace dr;

extern "C" {

// Replaces `dec_point` (dot_ring/vrf/codec.py:39-45) / `CurvePoint.string_to_point` (curve/point.py:178-214)
// for a batch: ok[i] = 0 <=> the reference raises ValueError.  checked != 0 adds the non-identity
// prime-subgroup test of `Curve.valid_point` (curve/curve.py:56-67).
int dr_te_decode_batch(dr_ctx* c, const uint8_t* in32, size_t n, int checked, uint8_t* out_xy64, uint8_t* ok) {
    try {
        Ctx* ctx = (Ctx*)c;
        if (!ctx || (n && (!in32 || !out_xy64 || !ok))) throw Error(DR_EINVAL, "bad argument");
        ctx->activate();
        if (!n) return DR_OK;
        DevBuf<uint8_t> din(n * 32), dout(n * 64), dok(n);
        h2d(ctx->stream, din.p, in32, n * 32);
        launch(ctx->stream, Dim3((uint32_t)((n + 63) / 64)), 64, 0, TeDecodeBody(), (const uint8_t*)din.p, (uint32_t)n, checked, dout.p, dok.p);
        d2h(ctx->stream, out_xy64, dout.p, n * 64);
        d2h(ctx->stream, ok, dok.p, n);
        stream_sync(ctx->stream);
    } catch (const Error& e) {
        return set_error(e.code, e.what());
    } catch (const std::exception& e) {
        return set_error(DR_ECUDA, e.what());
    }
    return DR_OK;
}

// Replaces `BandersnatchPoint.__mul__` (dot_ring/curve/specs/bandersnatch.py:177-191) for a batch:
// out[i] = (scalars[i] mod order) * points[i]  (or * points[0] when n_points == 1), 32-byte encodings.
int dr_te_mul_batch(dr_ctx* c, const uint8_t* points32, size_t n_points, const uint8_t* scalars32, size_t n, uint8_t* out32, uint8_t* ok) {
    try {
        Ctx* ctx = (Ctx*)c;
        if (!ctx || (n && (!points32 || !scalars32 || !out32 || !ok)) || (n_points != 1 && n_points != n)) throw Error(DR_EINVAL, "bad argument");
        ctx->activate();
        if (!n) return DR_OK;
        DevBuf<uint8_t> dp(n_points * 32), dk(n * 32), dout(n * 32), dok(n);
        h2d(ctx->stream, dp.p, points32, n_points * 32);
        h2d(ctx->stream, dk.p, scalars32, n * 32);
        launch(ctx->stream, Dim3((uint32_t)((n + 63) / 64)), 64, 0, TeMulBody(), (const uint8_t*)dp.p, (uint32_t)n_points, (const uint8_t*)dk.p, (uint32_t)n, dout.p, dok.p);
        d2h(ctx->stream, out32, dout.p, n * 32);
        d2h(ctx->stream, ok, dok.p, n);
        stream_sync(ctx->stream);
    } catch (const Error& e) {
        return set_error(e.code, e.what());
    } catch (const std::exception& e) {
        return set_error(DR_ECUDA, e.what());
    }
    return DR_OK;
}
}
