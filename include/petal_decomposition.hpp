// petal_decomposition.hpp - header-only C++ host mirror of petal-decomposition's public API
// (reference src/lib.rs:17-28) over the C ABI of libpetal_b200.so (include/petal_b200.h).
//
// Same type / method names and argument meaning as the Rust crate; `A` is float or double;
// matrices are row-major (samples x features) std::vector<A> + shape, host memory.  Errors map to
// the reference's DecompositionError::{InvalidInput, LinalgError} (src/lib.rs:22-28) as exceptions.
// This is the host side a C++ caller links against; the Rust crate keeps its own types and calls
// the same C ABI through the shim shown in INTEGRATION.md / rust/.
#pragma once
#include <cstdint>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "petal_b200.h"

namespace petal_decomposition {

struct DecompositionError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct InvalidInput : DecompositionError {  // src/lib.rs:24-25
    explicit InvalidInput(const std::string& m) : DecompositionError("invalid matrix: " + m) {}
};
struct LinalgError : DecompositionError {  // src/lib.rs:26-27 (sic)
    explicit LinalgError(const std::string& m) : DecompositionError("linear algerba operation failed: " + m) {}
};

// Row-major matrix, the stand-in for ndarray::Array2<A> in standard layout (src/linalg.rs:75).
template <typename A>
struct Matrix {
    int64_t rows = 0, cols = 0;
    std::vector<A> data;
    Matrix() {}
    Matrix(int64_t r, int64_t c) : rows(r), cols(c), data((size_t)(r * c)) {}
    A& operator()(int64_t i, int64_t j) { return data[(size_t)(i * cols + j)]; }
    const A& operator()(int64_t i, int64_t j) const { return data[(size_t)(i * cols + j)]; }
};

class Context {
public:
    explicit Context(int device = 0) {
        if (petal_ctx_create(device, &ctx_) != PETAL_OK) throw LinalgError(petal_last_global_error());
    }
    ~Context() { petal_ctx_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    petal_ctx* get() const { return ctx_; }
    void check(int status) const {
        if (status == PETAL_OK) return;
        if (status == PETAL_INVALID_INPUT) throw InvalidInput(petal_last_error(ctx_));
        throw LinalgError(petal_last_error(ctx_));
    }
    static std::shared_ptr<Context> shared() {
        static std::shared_ptr<Context> c = std::make_shared<Context>(0);
        return c;
    }
    // How a host Matrix<A> reaches HBM (petal_b200.h): 0 resident copy when it fits, else out of core; 1 resident;
    // 2 out of core (two ring slots).  chunk_bytes <= 0 keeps the current H2D chunk size.
    int set_host_staging(int mode, int64_t chunk_bytes = 0) { return petal_ctx_set_host_staging(ctx_, mode, chunk_bytes); }
    // Randomized PCA on host data: power iterations on the Gram matrix accumulated during the ingest (default on).
    int set_host_gram(int enable) { return petal_ctx_set_host_gram(ctx_, enable); }

private:
    petal_ctx* ctx_ = nullptr;
};

// rand_pcg::Mcg128Xsl64 as used by the reference (src/pca.rs:11-12, src/ica.rs:10-11).
class Pcg {
public:
    static Pcg from_seed(unsigned __int128 seed) {  // Pcg::from_seed(seed.to_be_bytes())
        return Pcg(petal_rng_from_seed((uint64_t)(seed >> 64), (uint64_t)seed));
    }
    static Pcg from_state(unsigned __int128 state) {  // Pcg64Mcg::new(state)
        return Pcg(petal_rng_from_state((uint64_t)(state >> 64), (uint64_t)state));
    }
    static Pcg from_entropy() {
        std::random_device rd;
        unsigned __int128 s = 0;
        for (int i = 0; i < 4; ++i) s = (s << 32) | rd();
        return from_seed(s);
    }
    Pcg(Pcg&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    Pcg(const Pcg&) = delete;
    ~Pcg() {
        if (h_) petal_rng_free(h_);
    }
    template <typename A>
    Matrix<A> standard_normal(int64_t rows, int64_t cols) {  // from_shape_fn(.., StandardNormal), row-major
        Matrix<A> m(rows, cols);
        if constexpr (sizeof(A) == 4) petal_rng_normal_f32(h_, (float*)m.data.data(), rows * cols);
        else petal_rng_normal_f64(h_, (double*)m.data.data(), rows * cols);
        return m;
    }

private:
    explicit Pcg(petal_rng* h) : h_(h) {}
    petal_rng* h_;
};

namespace detail {
template <typename A>
struct Abi;
template <>
struct Abi<float> {
    static constexpr auto pca_fit = petal_pca_fit_f32;
    static constexpr auto rpca_fit = petal_rpca_fit_f32;
    static constexpr auto transform = petal_transform_f32;
    static constexpr auto inverse_transform = petal_inverse_transform_f32;
    static constexpr auto fastica_fit = petal_fastica_fit_f32;
    static constexpr auto fastica_deflation_fit = petal_fastica_deflation_fit_f32;
};
template <>
struct Abi<double> {
    static constexpr auto pca_fit = petal_pca_fit_f64;
    static constexpr auto rpca_fit = petal_rpca_fit_f64;
    static constexpr auto transform = petal_transform_f64;
    static constexpr auto inverse_transform = petal_inverse_transform_f64;
    static constexpr auto fastica_fit = petal_fastica_fit_f64;
    static constexpr auto fastica_deflation_fit = petal_fastica_deflation_fit_f64;
};

template <typename A>
Matrix<A> transform(const Context& c, const Matrix<A>& x, const Matrix<A>& comps, const std::vector<A>& means,
                    bool centering) {  // src/pca.rs:726-750
    if (x.cols != (int64_t)means.size()) throw InvalidInput("# of columns should be " + std::to_string(means.size()));
    Matrix<A> out(x.rows, comps.rows);
    c.check(Abi<A>::transform(c.get(), x.data.data(), x.rows, x.cols, comps.data.data(), comps.rows,
                              centering ? means.data() : nullptr, out.data.data()));
    return out;
}
template <typename A>
Matrix<A> inverse_transform(const Context& c, const Matrix<A>& y, const Matrix<A>& comps, const std::vector<A>& means,
                            bool centering) {  // src/pca.rs:788-811
    if (y.cols != comps.rows) throw InvalidInput("# of columns should be " + std::to_string(comps.rows));
    Matrix<A> out(y.rows, comps.cols);
    c.check(Abi<A>::inverse_transform(c.get(), y.data.data(), y.rows, y.cols, comps.data.data(), comps.cols,
                                      centering ? means.data() : nullptr, out.data.data()));
    return out;
}
}  // namespace detail

// ---------------------------------------------------------------------------------------------
// Pca<A> / PcaBuilder (src/pca.rs:41-283)
// ---------------------------------------------------------------------------------------------
template <typename A>
class Pca {
public:
    explicit Pca(int64_t n_components, bool centering = true, std::shared_ptr<Context> ctx = Context::shared())
        : k_(n_components), centering_(centering), ctx_(std::move(ctx)) {
        components_.rows = n_components;  // Array2::zeros((n_components, 0))
    }
    const Matrix<A>& components() const { return components_; }
    const std::vector<A>& mean() const { return means_; }
    int64_t n_components() const { return components_.rows; }
    const std::vector<A>& singular_values() const { return singular_; }
    std::vector<A> explained_variance_ratio() const {  // src/pca.rs:101-105
        std::vector<A> r(singular_.size());
        for (size_t i = 0; i < r.size(); ++i) r[i] = singular_[i] * singular_[i] / total_variance_;
        return r;
    }
    void fit(const Matrix<A>& x) { inner_fit(x, nullptr); }
    Matrix<A> fit_transform(const Matrix<A>& x) {
        Matrix<A> y(x.rows, k_);
        inner_fit(x, &y);
        return y;
    }
    Matrix<A> transform(const Matrix<A>& x) const { return detail::transform(*ctx_, x, components_, means_, centering_); }
    Matrix<A> inverse_transform(const Matrix<A>& y) const {
        return detail::inverse_transform(*ctx_, y, components_, means_, centering_);
    }

private:
    void inner_fit(const Matrix<A>& x, Matrix<A>* scores) {  // src/pca.rs:195-231
        if (x.rows < k_ || x.cols < k_) throw InvalidInput("every dimension should be at least " + std::to_string(k_));
        if (x.rows == 0) return;  // mean_axis -> None, src/pca.rs:207-211
        Matrix<A> comps(k_, x.cols);
        std::vector<A> mean((size_t)x.cols), sing((size_t)k_);
        A tv = 0;
        ctx_->check(detail::Abi<A>::pca_fit(ctx_->get(), x.data.data(), x.rows, x.cols, k_, centering_ ? 1 : 0,
                                            comps.data.data(), mean.data(), sing.data(), &tv,
                                            scores ? scores->data.data() : nullptr));
        components_ = std::move(comps);
        means_ = std::move(mean);
        singular_ = std::move(sing);
        total_variance_ = tv;
        n_samples_ = x.rows;
    }
    int64_t k_;
    bool centering_;
    std::shared_ptr<Context> ctx_;
    Matrix<A> components_;
    std::vector<A> means_, singular_;
    A total_variance_ = 0;
    int64_t n_samples_ = 0;
};

class PcaBuilder {  // src/pca.rs:246-283
public:
    explicit PcaBuilder(int64_t n_components) : k_(n_components) {}
    PcaBuilder& centering(bool c) {
        centering_ = c;
        return *this;
    }
    template <typename A>
    Pca<A> build() const {
        return Pca<A>(k_, centering_);
    }

private:
    int64_t k_;
    bool centering_ = true;
};

// ---------------------------------------------------------------------------------------------
// RandomizedPca<A> / RandomizedPcaBuilder (src/pca.rs:317-663)
// ---------------------------------------------------------------------------------------------
template <typename A>
class RandomizedPca {
public:
    RandomizedPca(int64_t n_components, Pcg rng, bool centering = true,
                  std::shared_ptr<Context> ctx = Context::shared())
        : k_(n_components), centering_(centering), rng_(std::move(rng)), ctx_(std::move(ctx)) {
        components_.rows = n_components;
    }
    static RandomizedPca with_seed(int64_t k, unsigned __int128 seed) { return RandomizedPca(k, Pcg::from_seed(seed)); }
    static RandomizedPca with_rng(int64_t k, Pcg rng) { return RandomizedPca(k, std::move(rng)); }  // src/pca.rs:371
    int64_t n_oversamples = 10;  // src/pca.rs:679
    int64_t n_power_iter = 7;    // src/pca.rs:680

    const Matrix<A>& components() const { return components_; }
    const std::vector<A>& mean() const { return means_; }
    int64_t n_components() const { return components_.rows; }
    const std::vector<A>& singular_values() const { return singular_; }
    std::vector<A> explained_variance_ratio() const {
        std::vector<A> r(singular_.size());
        for (size_t i = 0; i < r.size(); ++i) r[i] = singular_[i] * singular_[i] / total_variance_;
        return r;
    }
    void fit(const Matrix<A>& x) { inner_fit(x, nullptr); }
    Matrix<A> fit_transform(const Matrix<A>& x) {
        Matrix<A> y(x.rows, k_);
        inner_fit(x, &y);
        return y;
    }
    Matrix<A> transform(const Matrix<A>& x) const { return detail::transform(*ctx_, x, components_, means_, centering_); }
    Matrix<A> inverse_transform(const Matrix<A>& y) const {
        return detail::inverse_transform(*ctx_, y, components_, means_, centering_);
    }

private:
    void inner_fit(const Matrix<A>& x, Matrix<A>* scores) {  // src/pca.rs:509-550
        if (x.rows < k_ || x.cols < k_) throw InvalidInput("every dimension should be at least " + std::to_string(k_));
        if (x.rows == 0) return;
        // Omega: d x (k + oversamples), row-major draw order (src/pca.rs:701-705); advances the model's rng
        Matrix<A> omega = rng_.template standard_normal<A>(x.cols, k_ + n_oversamples);
        Matrix<A> comps(k_, x.cols);
        std::vector<A> mean((size_t)x.cols), sing((size_t)k_);
        A tv = 0;
        ctx_->check(detail::Abi<A>::rpca_fit(ctx_->get(), x.data.data(), x.rows, x.cols, k_, centering_ ? 1 : 0,
                                             n_oversamples, n_power_iter, omega.data.data(), comps.data.data(),
                                             mean.data(), sing.data(), &tv, scores ? scores->data.data() : nullptr));
        components_ = std::move(comps);
        means_ = std::move(mean);
        singular_ = std::move(sing);
        total_variance_ = tv;
    }
    int64_t k_;
    bool centering_;
    Pcg rng_;
    std::shared_ptr<Context> ctx_;
    Matrix<A> components_;
    std::vector<A> means_, singular_;
    A total_variance_ = 0;
};

class RandomizedPcaBuilder {  // src/pca.rs:564-663
public:
    explicit RandomizedPcaBuilder(int64_t n_components) : k_(n_components) {}
    RandomizedPcaBuilder& seed(unsigned __int128 s) {
        seed_ = s;
        seeded_ = true;
        return *this;
    }
    RandomizedPcaBuilder& centering(bool c) {
        centering_ = c;
        return *this;
    }
    template <typename A>
    RandomizedPca<A> build() const {
        return RandomizedPca<A>(k_, seeded_ ? Pcg::from_seed(seed_) : Pcg::from_entropy(), centering_);
    }

private:
    int64_t k_;
    unsigned __int128 seed_ = 0;
    bool seeded_ = false, centering_ = true;
};

// ---------------------------------------------------------------------------------------------
// FastIca<A> / FastIcaBuilder (src/ica.rs:41-317): fit, transform, fit_transform only
// ---------------------------------------------------------------------------------------------
template <typename A>
class FastIca {
public:
    explicit FastIca(Pcg rng, std::shared_ptr<Context> ctx = Context::shared()) : rng_(std::move(rng)), ctx_(std::move(ctx)) {}
    static FastIca with_seed(unsigned __int128 seed) { return FastIca(Pcg::from_seed(seed)); }
    static FastIca with_rng(Pcg rng) { return FastIca(std::move(rng)); }
    void fit(const Matrix<A>& x) { inner_fit(x, nullptr); }
    Matrix<A> fit_transform(const Matrix<A>& x) {
        Matrix<A> s(x.rows, std::min(x.rows, x.cols));
        inner_fit(x, &s);
        return s;
    }
    Matrix<A> transform(const Matrix<A>& x) const {
        if (x.cols != (int64_t)means_.size()) throw InvalidInput("too many columns");  // src/ica.rs:124-128
        return detail::transform(*ctx_, x, components_, means_, true);
    }
    int64_t n_iter = 0;  // private field in the reference, read by its tests (src/ica.rs:412)
    bool deflation = false;  // extension: deflation scheme instead of the reference's symmetric one

private:
    void inner_fit(const Matrix<A>& x, Matrix<A>* sources) {  // src/ica.rs:167-222
        if (x.rows == 0) return;
        const int64_t nc = std::min(x.rows, x.cols);
        Matrix<A> w_init = rng_.template standard_normal<A>(nc, nc);  // src/ica.rs:210-214
        Matrix<A> comps(nc, x.cols);
        std::vector<A> mean((size_t)x.cols);
        double lim = 0;
        if (deflation)  // extension: one component at a time; the reference has the symmetric scheme only
            ctx_->check(detail::Abi<A>::fastica_deflation_fit(ctx_->get(), x.data.data(), x.rows, x.cols, PETAL_ICA_LOGCOSH,
                                                              1e-4, 200, w_init.data.data(), comps.data.data(), mean.data(),
                                                              &n_iter, &lim, sources ? sources->data.data() : nullptr));
        else
            ctx_->check(detail::Abi<A>::fastica_fit(ctx_->get(), x.data.data(), x.rows, x.cols, PETAL_ICA_LOGCOSH, 1e-4, 200,
                                                    0, w_init.data.data(), comps.data.data(), mean.data(), &n_iter, &lim,
                                                    sources ? sources->data.data() : nullptr));
        components_ = std::move(comps);
        means_ = std::move(mean);
    }
    Pcg rng_;
    std::shared_ptr<Context> ctx_;
    Matrix<A> components_;
    std::vector<A> means_;
};

class FastIcaBuilder {  // src/ica.rs:244-317
public:
    FastIcaBuilder& seed(unsigned __int128 s) {
        seed_ = s;
        seeded_ = true;
        return *this;
    }
    template <typename A>
    FastIca<A> build() const {
        return FastIca<A>(seeded_ ? Pcg::from_seed(seed_) : Pcg::from_entropy());
    }

private:
    unsigned __int128 seed_ = 0;
    bool seeded_ = false;
};

}  // namespace petal_decomposition
