// Host-compiled view of the device math headers (portable branch), for CPU-side unit tests.
#include "ge25519.cuh"
#include "sc25519.cuh"
#include <string.h>
extern "C" {
// scalar field: built twice by the test (4 x 64-bit host limbs, and -DBPG_SC_PORTABLE = the 8 x 32-bit code the device runs)
void hm_sc_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) { sc x, y; memcpy(&x, a, 32); memcpy(&y, b, 32); sc z = sc_mul(x, y); memcpy(r, &z, 32); }
void hm_sc_montmul(const uint32_t* a, const uint32_t* b, uint32_t* r) { sc x, y; memcpy(&x, a, 32); memcpy(&y, b, 32); sc z = sc_montmul(x, y); memcpy(r, &z, 32); }
void hm_sc_add(const uint32_t* a, const uint32_t* b, uint32_t* r) { sc x, y; memcpy(&x, a, 32); memcpy(&y, b, 32); sc z = sc_add(x, y); memcpy(r, &z, 32); }
void hm_sc_sub(const uint32_t* a, const uint32_t* b, uint32_t* r) { sc x, y; memcpy(&x, a, 32); memcpy(&y, b, 32); sc z = sc_sub(x, y); memcpy(r, &z, 32); }
void hm_sc_reduce(const uint32_t* a, uint32_t* r) { sc x; memcpy(&x, a, 32); sc z = sc_reduce(x); memcpy(r, &z, 32); }
uint32_t hm_sc_add_raw(const uint32_t* a, const uint32_t* b, uint32_t* r) { sc x, y, z; memcpy(&x, a, 32); memcpy(&y, b, 32); uint32_t c = sc_add_raw(&z, x, y); memcpy(r, &z, 32); return c; }
uint32_t hm_sc_sub_raw(const uint32_t* a, const uint32_t* b, uint32_t* r) { sc x, y, z; memcpy(&x, a, 32); memcpy(&y, b, 32); uint32_t c = sc_sub_raw(&z, x, y); memcpy(r, &z, 32); return c; }
void hm_fe_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) { fe x, y; memcpy(&x, a, 32); memcpy(&y, b, 32); fe z = fe_mul(x, y); memcpy(r, &z, 32); }
void hm_fe_add(const uint32_t* a, const uint32_t* b, uint32_t* r) { fe x, y; memcpy(&x, a, 32); memcpy(&y, b, 32); fe z = fe_add(x, y); memcpy(r, &z, 32); }
void hm_fe_sub(const uint32_t* a, const uint32_t* b, uint32_t* r) { fe x, y; memcpy(&x, a, 32); memcpy(&y, b, 32); fe z = fe_sub(x, y); memcpy(r, &z, 32); }
void hm_fe_canon(const uint32_t* a, uint32_t* r) { fe x; memcpy(&x, a, 32); fe z = fe_canon(x); memcpy(r, &z, 32); }
void hm_fe_invert(const uint32_t* a, uint32_t* r) { fe x; memcpy(&x, a, 32); fe z = fe_canon(fe_invert(x)); memcpy(r, &z, 32); }
void hm_from_uniform(const uint8_t* b64, uint8_t* out32) { ge_ext p = ge_from_uniform_bytes(b64); ge_ristretto_compress(out32, p); }
int hm_decompress_ops(const uint8_t* in32, uint8_t* out_same, uint8_t* out_dbl, uint8_t* out_niels_rt) {
    ge_ext p; if (!ge_ristretto_decompress(&p, in32)) return 0;
    ge_ristretto_compress(out_same, p);
    ge_ext d = ge_dbl(p); ge_ristretto_compress(out_dbl, d);
    // niels round trip: (d -> niels) added to p, minus p  == d
    ge_niels n = ge_to_niels(d, fe_invert(d.Z));
    ge_ext q = ge_madd(p, n, false); ge_niels pn = ge_to_niels(p, fe_invert(p.Z)); q = ge_madd(q, pn, true);
    ge_ristretto_compress(out_niels_rt, q);
    return 1;
}
int hm_add(const uint8_t* a32, const uint8_t* b32, uint8_t* out) { ge_ext p, q; if (!ge_ristretto_decompress(&p, a32) || !ge_ristretto_decompress(&q, b32)) return 0; ge_ristretto_compress(out, ge_add(p, q)); return 1; }
}
