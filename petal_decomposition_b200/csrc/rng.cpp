// Host RNG: the reference's seeded stream for Omega / w_init.
//
// rand_pcg::Mcg128Xsl64 (reference src/pca.rs:9-12,356-358; src/ica.rs:8-11,75-77) sampled
// through rand_distr::StandardNormal (src/pca.rs:701-705; src/ica.rs:210-214).  Neither crate
// is vendored in the reference tree, so the published algorithms are restated here:
// PCG XSL-RR 128/64 MCG and the 256-layer ziggurat (tables built with the recurrence rand's
// table generator uses: R = 3.6541528853610088, V = 0.00492867323399).
#include <cmath>
#include <cstdint>
#include <cstring>

#include "../../include/petal_b200.h"

namespace {

typedef unsigned __int128 u128;

const u128 kMult = ((u128)0x2360ED051FC65DA4ULL << 64) | (u128)0x4385DF649FCCF645ULL;
const double kZigR = 3.6541528853610088;
const double kZigV = 0.00492867323399;

struct ZigTables {
    double x[257];
    double f[257];
    ZigTables() {
        auto pdf = [](double v) { return std::exp(-v * v / 2.0); };
        x[0] = kZigV / pdf(kZigR);
        x[1] = kZigR;
        for (int i = 2; i < 256; ++i) {
            double last = x[i - 1];
            x[i] = std::sqrt(-2.0 * std::log(kZigV / last + pdf(last)));
        }
        x[256] = 0.0;
        for (int i = 0; i <= 256; ++i) f[i] = pdf(x[i]);
    }
};
const ZigTables& tables() {
    static ZigTables t;
    return t;
}

inline double bits_to_f64(uint64_t b) {
    double d;
    std::memcpy(&d, &b, sizeof d);
    return d;
}

}  // namespace

struct petal_rng {
    u128 state;
    uint64_t next() {
        state *= kMult;
        unsigned rot = (unsigned)(state >> 122);
        uint64_t xsl = (uint64_t)(state >> 64) ^ (uint64_t)state;
        return (xsl >> rot) | (xsl << ((64 - rot) & 63));
    }
    double standard() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    double open01() { return bits_to_f64((next() >> 12) | (1023ULL << 52)) - (1.0 - 0x1p-53); }
    double tail(double u) {
        double x = 1.0, y = 0.0;
        while (-2.0 * y < x * x) {
            double x_ = open01();
            double y_ = open01();
            x = std::log(x_) / kZigR;
            y = std::log(y_);
        }
        return u < 0.0 ? x - kZigR : kZigR - x;
    }
    double normal() {
        const ZigTables& t = tables();
        for (;;) {
            uint64_t bits = next();
            unsigned i = (unsigned)(bits & 0xff);
            double u = bits_to_f64((bits >> 12) | (1024ULL << 52)) - 3.0;
            double x = u * t.x[i];
            if (std::fabs(x) < t.x[i + 1]) return x;
            if (i == 0) return tail(u);
            if (t.f[i + 1] + (t.f[i] - t.f[i + 1]) * standard() < std::exp(-x * x / 2.0)) return x;
        }
    }
};

extern "C" {

petal_rng* petal_rng_from_state(uint64_t state_hi, uint64_t state_lo) {
    petal_rng* r = new petal_rng;
    r->state = (((u128)state_hi << 64) | (u128)state_lo) | 1;
    return r;
}

petal_rng* petal_rng_from_seed(uint64_t seed_hi, uint64_t seed_lo) {
    // from_seed(seed.to_be_bytes()) reads the big-endian bytes as a little-endian u128:
    // the state is the byte-swapped seed.
    return petal_rng_from_state(__builtin_bswap64(seed_lo), __builtin_bswap64(seed_hi));
}

void petal_rng_free(petal_rng* rng) { delete rng; }

uint64_t petal_rng_next_u64(petal_rng* rng) { return rng->next(); }

void petal_rng_get_state(const petal_rng* rng, uint64_t* state_hi, uint64_t* state_lo) {
    *state_hi = (uint64_t)(rng->state >> 64);
    *state_lo = (uint64_t)rng->state;
}

void petal_rng_normal_f64(petal_rng* rng, double* out, int64_t count) {
    for (int64_t i = 0; i < count; ++i) out[i] = rng->normal();
}

void petal_rng_normal_f32(petal_rng* rng, float* out, int64_t count) {
    for (int64_t i = 0; i < count; ++i) out[i] = (float)rng->normal();
}

}  // extern "C"
