/* ORACLE -- test infrastructure only. Never linked into the product.
 *
 * Short-Weierstrass y^2 = x^3 + 5 in Jacobian coordinates (x = X/Z^2, y = Y/Z^3), restating
 * the group law semantics of pasta_curves 0.4.1 `curves.rs` (Cargo.lock:567-579; not vendored):
 * affine identity = (0,0); compressed = x LE | y-parity<<255; identity = 32 zero bytes.
 * Group-element outputs are canonicalised through to_affine, so the particular addition
 * formulas are free.
 *
 * Include with: #define EC(name) vesta_##name, #define BF(name) fq_##name (base field),
 *               #define SF(name) fp_##name (scalar field)
 */

typedef struct { BF(t) x, y; } EC(affine);            /* identity: x = y = 0 */
typedef struct { BF(t) x, y, z; } EC(point);          /* identity: z = 0 */

static inline int EC(affine_is_identity)(const EC(affine)* a) { return BF(is_zero)(&a->x) && BF(is_zero)(&a->y); }
static inline int EC(is_identity)(const EC(point)* a) { return BF(is_zero)(&a->z); }
static inline void EC(set_identity)(EC(point)* r) { memset(r, 0, sizeof(*r)); r->y = BF(R1); }
static inline void EC(from_affine)(EC(point)* r, const EC(affine)* a) {
  if (EC(affine_is_identity)(a)) { EC(set_identity)(r); return; }
  r->x = a->x; r->y = a->y; r->z = BF(R1);
}

static inline void EC(dbl)(EC(point)* r, const EC(point)* p) {
  if (EC(is_identity)(p)) { *r = *p; return; }
  BF(t) a, b, c, d, e, f, t, x3, y3, z3;
  BF(sqr)(&a, &p->x);
  BF(sqr)(&b, &p->y);
  BF(sqr)(&c, &b);
  BF(add)(&t, &p->x, &b); BF(sqr)(&t, &t); BF(sub)(&t, &t, &a); BF(sub)(&t, &t, &c); BF(dbl)(&d, &t);
  BF(dbl)(&e, &a); BF(add)(&e, &e, &a);
  BF(sqr)(&f, &e);
  BF(dbl)(&t, &d); BF(sub)(&x3, &f, &t);
  BF(sub)(&t, &d, &x3); BF(mul)(&y3, &e, &t);
  BF(dbl)(&t, &c); BF(dbl)(&t, &t); BF(dbl)(&t, &t); BF(sub)(&y3, &y3, &t);
  BF(mul)(&z3, &p->y, &p->z); BF(dbl)(&z3, &z3);
  r->x = x3; r->y = y3; r->z = z3;
}

static inline void EC(add)(EC(point)* r, const EC(point)* p, const EC(point)* q) {
  if (EC(is_identity)(p)) { *r = *q; return; }
  if (EC(is_identity)(q)) { *r = *p; return; }
  BF(t) z1z1, z2z2, u1, u2, s1, s2, h, i, j, rr, v, t, x3, y3, z3;
  BF(sqr)(&z1z1, &p->z); BF(sqr)(&z2z2, &q->z);
  BF(mul)(&u1, &p->x, &z2z2); BF(mul)(&u2, &q->x, &z1z1);
  BF(mul)(&s1, &p->y, &q->z); BF(mul)(&s1, &s1, &z2z2);
  BF(mul)(&s2, &q->y, &p->z); BF(mul)(&s2, &s2, &z1z1);
  if (BF(eq)(&u1, &u2)) {
    if (BF(eq)(&s1, &s2)) { EC(dbl)(r, p); return; }
    EC(set_identity)(r); return;
  }
  BF(sub)(&h, &u2, &u1);
  BF(dbl)(&i, &h); BF(sqr)(&i, &i);
  BF(mul)(&j, &h, &i);
  BF(sub)(&rr, &s2, &s1); BF(dbl)(&rr, &rr);
  BF(mul)(&v, &u1, &i);
  BF(sqr)(&x3, &rr); BF(sub)(&x3, &x3, &j); BF(dbl)(&t, &v); BF(sub)(&x3, &x3, &t);
  BF(sub)(&t, &v, &x3); BF(mul)(&y3, &rr, &t); BF(mul)(&t, &s1, &j); BF(dbl)(&t, &t); BF(sub)(&y3, &y3, &t);
  BF(add)(&z3, &p->z, &q->z); BF(sqr)(&z3, &z3); BF(sub)(&z3, &z3, &z1z1); BF(sub)(&z3, &z3, &z2z2); BF(mul)(&z3, &z3, &h);
  r->x = x3; r->y = y3; r->z = z3;
}

static inline void EC(add_mixed)(EC(point)* r, const EC(point)* p, const EC(affine)* q) {
  if (EC(affine_is_identity)(q)) { *r = *p; return; }
  if (EC(is_identity)(p)) { EC(from_affine)(r, q); return; }
  BF(t) z1z1, u2, s2, h, hh, i, j, rr, v, t, x3, y3, z3;
  BF(sqr)(&z1z1, &p->z);
  BF(mul)(&u2, &q->x, &z1z1);
  BF(mul)(&s2, &q->y, &p->z); BF(mul)(&s2, &s2, &z1z1);
  if (BF(eq)(&p->x, &u2)) {
    if (BF(eq)(&p->y, &s2)) { EC(dbl)(r, p); return; }
    EC(set_identity)(r); return;
  }
  BF(sub)(&h, &u2, &p->x);
  BF(sqr)(&hh, &h);
  BF(dbl)(&i, &hh); BF(dbl)(&i, &i);
  BF(mul)(&j, &h, &i);
  BF(sub)(&rr, &s2, &p->y); BF(dbl)(&rr, &rr);
  BF(mul)(&v, &p->x, &i);
  BF(sqr)(&x3, &rr); BF(sub)(&x3, &x3, &j); BF(dbl)(&t, &v); BF(sub)(&x3, &x3, &t);
  BF(sub)(&t, &v, &x3); BF(mul)(&y3, &rr, &t); BF(mul)(&t, &p->y, &j); BF(dbl)(&t, &t); BF(sub)(&y3, &y3, &t);
  BF(add)(&z3, &p->z, &h); BF(sqr)(&z3, &z3); BF(sub)(&z3, &z3, &z1z1); BF(sub)(&z3, &z3, &hh);
  r->x = x3; r->y = y3; r->z = z3;
}

static inline void EC(neg)(EC(point)* r, const EC(point)* p) { r->x = p->x; BF(neg)(&r->y, &p->y); r->z = p->z; }
static inline void EC(sub)(EC(point)* r, const EC(point)* p, const EC(point)* q) { EC(point) n; EC(neg)(&n, q); EC(add)(r, p, &n); }

static inline void EC(to_affine)(EC(affine)* r, const EC(point)* p) {
  if (EC(is_identity)(p)) { memset(r, 0, sizeof(*r)); return; }
  BF(t) zi, zi2, zi3;
  BF(inv)(&zi, &p->z); BF(sqr)(&zi2, &zi); BF(mul)(&zi3, &zi2, &zi);
  BF(mul)(&r->x, &p->x, &zi2); BF(mul)(&r->y, &p->y, &zi3);
}

/* group::Curve::batch_normalize */
static inline void EC(batch_normalize)(EC(affine)* out, const EC(point)* in, size_t n) {
  BF(t)* zs = (BF(t)*)malloc(sizeof(BF(t)) * n * 2);
  for (size_t i = 0; i < n; ++i) zs[i] = in[i].z;
  BF(batch_invert)(zs, n, zs + n);
  for (size_t i = 0; i < n; ++i) {
    if (EC(is_identity)(&in[i])) { memset(&out[i], 0, sizeof(out[i])); continue; }
    BF(t) zi2, zi3;
    BF(sqr)(&zi2, &zs[i]); BF(mul)(&zi3, &zi2, &zs[i]);
    BF(mul)(&out[i].x, &in[i].x, &zi2); BF(mul)(&out[i].y, &in[i].y, &zi3);
  }
  free(zs);
}

/* scalar multiplication by a scalar-field element (double-and-add on canonical bits) */
static inline void EC(mul)(EC(point)* r, const EC(point)* p, const SF(t)* k) {
  uint64_t e[4]; SF(to_raw)(e, k);
  EC(point) acc; EC(set_identity)(&acc);
  int started = 0;
  for (int i = 255; i >= 0; --i) {
    if (started) EC(dbl)(&acc, &acc);
    if ((e[i >> 6] >> (i & 63)) & 1) { EC(add)(&acc, &acc, p); started = 1; }
  }
  *r = acc;
}

static inline int EC(on_curve_affine)(const EC(affine)* a) {
  if (EC(affine_is_identity)(a)) return 1;
  BF(t) l, r, five; BF(sqr)(&l, &a->y); BF(sqr)(&r, &a->x); BF(mul)(&r, &r, &a->x);
  BF(from_u64)(&five, 5); BF(add)(&r, &r, &five);
  return BF(eq)(&l, &r);
}

/* compressed encoding (32 B) */
static inline void EC(to_bytes)(uint8_t out[32], const EC(affine)* a) {
  if (EC(affine_is_identity)(a)) { memset(out, 0, 32); return; }
  uint8_t yb[32]; BF(to_repr)(out, &a->x); BF(to_repr)(yb, &a->y);
  out[31] |= (uint8_t)((yb[0] & 1) << 7);
}
