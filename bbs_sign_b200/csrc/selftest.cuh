// Self-test kernels behind bbs_selftest_* (include/bbs_b200.h): they expose the field, G1 and pairing
// layers one level below the BBS operations so the parity tests can pin each layer against the oracle.
// Included by capi.cu inside its anonymous namespace.

struct FieldTestArgs { int op; const uint8_t* a; const uint8_t* b; uint8_t* out; };

template <class C> BBS_HD void field_test_item(const FieldTestArgs& t, uint32_t i) {
    using Fp = typename C::Fp;
    using Fr = typename C::Fr;
    if (t.op <= 4) {
        constexpr int N = Fp::N;
        uint32_t a[N], b[N], r[N];
        limbs_from_le<N>(a, t.a + (size_t)i * 4 * N);
        limbs_from_le<N>(b, t.b + (size_t)i * 4 * N);
        fe_to_mont<Fp>(a, a); fe_to_mont<Fp>(b, b);
        if (t.op == 0) fe_mul<Fp>(r, a, b);
        else if (t.op == 1) fe_add<Fp>(r, a, b);
        else if (t.op == 2) fe_sub<Fp>(r, a, b);
        else if (t.op == 3) fe_inv<Fp>(r, a);
        else { if (!fe_sqrt<Fp>(r, a)) bn_zero<N>(r); }
        fe_from_mont<Fp>(r, r);
        limbs_to_le<N>(t.out + (size_t)i * 4 * N, r);
    } else {
        uint32_t a[8], b[8], r[8];
        limbs_from_le<8>(a, t.a + (size_t)i * 32);
        limbs_from_le<8>(b, t.b + (size_t)i * 32);
        fe_to_mont<Fr>(a, a); fe_to_mont<Fr>(b, b);
        if (t.op == 5) fe_mul<Fr>(r, a, b); else fe_inv<Fr>(r, a);
        fe_from_mont<Fr>(r, r);
        limbs_to_le<8>(t.out + (size_t)i * 32, r);
    }
}

template <class C> int selftest_field(int op, size_t n, const uint8_t* a, const uint8_t* b, uint8_t* out) {
    if (op < 0 || op > 6 || !a || !b || !out) return arg_error("selftest_field");
    const size_t w = op <= 4 ? 4 * C::Fp::N : 32;
    DevBuf da, db, dout;
    int rc = 0;
    if (!(rc = da.reserve(n * w)) && !(rc = db.reserve(n * w)) && !(rc = dout.reserve(n * w)) &&
        !(rc = rt_h2d(da.p, a, n * w, 0)) && !(rc = rt_h2d(db.p, b, n * w, 0))) {
        FieldTestArgs t{op, (const uint8_t*)da.p, (const uint8_t*)db.p, (uint8_t*)dout.p};
        rc = rt_launch<FieldTestArgs, &field_test_item<C>, 128>(t, (uint32_t)n, 0);
        if (!rc) rc = rt_d2h(out, dout.p, n * w, 0);
        if (!rc) rc = rt_sync(0);
    }
    da.release(); db.release(); dout.release();
    return rc;
}

struct G1MulTestArgs { const uint8_t* pts; const uint8_t* sc; uint8_t* out; };
template <class C> BBS_HD void g1_mul_test_item(const G1MulTestArgs& t, uint32_t i) {
    uint32_t A[G1A], k[8], R[G1J];
    uint8_t* o = t.out + (size_t)i * C::G1_BYTES;
    int st = g1_decompress<C>(A, t.pts + (size_t)i * C::G1_BYTES);
    limbs_from_le<8>(k, t.sc + (size_t)i * 32);
    if (st == PT_BAD) { for (int j = 0; j < C::G1_BYTES; j++) o[j] = 0xff; return; }
    if (st == PT_INF) g1_set_inf<C>(R); else g1_mul_affine<C>(R, A, k, 256);
    g1_compress<C>(o, R);
}
template <class C> int selftest_g1_mul(size_t n, const uint8_t* pts, const uint8_t* sc, uint8_t* out) {
    if (!pts || !sc || !out) return arg_error("selftest_g1_mul");
    DevBuf dp, ds, dout;
    int rc = 0;
    if (!(rc = dp.reserve(n * C::G1_BYTES)) && !(rc = ds.reserve(n * 32)) && !(rc = dout.reserve(n * C::G1_BYTES)) &&
        !(rc = rt_h2d(dp.p, pts, n * C::G1_BYTES, 0)) && !(rc = rt_h2d(ds.p, sc, n * 32, 0))) {
        G1MulTestArgs t{(const uint8_t*)dp.p, (const uint8_t*)ds.p, (uint8_t*)dout.p};
        rc = rt_launch<G1MulTestArgs, &g1_mul_test_item<C>, 128>(t, (uint32_t)n, 0);
        if (!rc) rc = rt_d2h(out, dout.p, n * C::G1_BYTES, 0);
        if (!rc) rc = rt_sync(0);
    }
    dp.release(); ds.release(); dout.release();
    return rc;
}

struct PairTestPrepArgs { const uint8_t* q; uint32_t* W; uint32_t* lines; uint32_t* st; };
template <class C> BBS_HD void pair_test_prep_item(const PairTestPrepArgs& t, uint32_t i) {
    if (i == 0) {
        int s = g2_decompress<C>(t.W, t.q);
        t.st[0] = (uint32_t)s;
        if (s == PT_OK) g2_precompute_lines<C>(t.lines, t.W, 0);
    } else {
        g2_precompute_lines<C>(t.lines, C::G2(), 1);
    }
}
struct PairTestArgs { const uint8_t* p; const uint8_t* r; const uint32_t* lines; const uint32_t* st; uint8_t* status; };
template <class C> BBS_HD void pair_test_item(const PairTestArgs& t, uint32_t i) {
    uint32_t P[G1A], R[G1A], a0[3 * FPN], a1[3 * FPN], f[F12N];
    int sp = g1_decompress<C>(P, t.p + (size_t)i * C::G1_BYTES);
    int sr = g1_decompress<C>(R, t.r + (size_t)i * C::G1_BYTES);
    if (sp == PT_BAD || sr == PT_BAD || t.st[0] == PT_BAD) { t.status[i] = ST_ERR_MALFORMED; return; }
    bn_copy<2 * C::Fp::N>(a0, P); fe_set_one<typename C::Fp>(a0 + 2 * FPN);
    bn_copy<2 * C::Fp::N>(a1, R); fe_set_one<typename C::Fp>(a1 + 2 * FPN);
    miller2<C>(f, t.lines, a0, sp == PT_INF || t.st[0] == PT_INF, a1, sr == PT_INF);
    t.status[i] = final_exp_is_one<C>(f) ? ST_ACCEPT : ST_REJECT;
}
template <class C> int selftest_pairing(size_t n, const uint8_t* p_points, const uint8_t* r_points, const uint8_t* q_point,
                                        uint8_t* status) {
    if (!p_points || !r_points || !q_point || !status) return arg_error("selftest_pairing");
    DevBuf dp, dr, dq, dW, dl, dst, dout;
    const size_t line_bytes = (size_t)ate_line_count<C>() * 2 * 4 * C::Fp::N * 4;
    int rc = 0;
    if (!(rc = dp.reserve(n * C::G1_BYTES)) && !(rc = dr.reserve(n * C::G1_BYTES)) && !(rc = dq.reserve(C::G2_BYTES)) &&
        !(rc = dW.reserve(4 * C::Fp::N * 4)) && !(rc = dl.reserve(line_bytes)) && !(rc = dst.reserve(16)) &&
        !(rc = dout.reserve(n)) && !(rc = rt_h2d(dp.p, p_points, n * C::G1_BYTES, 0)) &&
        !(rc = rt_h2d(dr.p, r_points, n * C::G1_BYTES, 0)) && !(rc = rt_h2d(dq.p, q_point, C::G2_BYTES, 0)) &&
        !(rc = rt_memset(dl.p, 0, line_bytes, 0))) {
        PairTestPrepArgs pa{(const uint8_t*)dq.p, (uint32_t*)dW.p, (uint32_t*)dl.p, (uint32_t*)dst.p};
        rc = rt_launch<PairTestPrepArgs, &pair_test_prep_item<C>, 32>(pa, 2, 0);
        PairTestArgs t{(const uint8_t*)dp.p, (const uint8_t*)dr.p, (const uint32_t*)dl.p, (const uint32_t*)dst.p,
                       (uint8_t*)dout.p};
        if (!rc) rc = rt_launch<PairTestArgs, &pair_test_item<C>, 128>(t, (uint32_t)n, 0);
        if (!rc) rc = rt_d2h(status, dout.p, n, 0);
        if (!rc) rc = rt_sync(0);
    }
    dp.release(); dr.release(); dq.release(); dW.release(); dl.release(); dst.release(); dout.release();
    return rc;
}
