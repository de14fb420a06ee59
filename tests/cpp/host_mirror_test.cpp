#include "petal_decomposition.hpp"
#include <cstdio>
#include <cmath>
using namespace petal_decomposition;
int main() {
    // reference test `pca` (src/pca.rs:886-906) replayed through the C++ host mirror
    try {
        Matrix<double> x(3, 2);
        double v[6] = {0, 0, 3, 4, 6, 8};
        for (int i = 0; i < 6; ++i) x.data[i] = v[i];
        Pca<double> pca(1);
        auto y = pca.fit_transform(x);
        std::printf("%g %g %g\n", std::fabs(y(0, 0)), y(1, 0), std::fabs(y(2, 0)));
        auto r = RandomizedPca<double>::with_seed(1, 1234567891011121314ULL);
        r.fit(x);
        auto ica = FastIcaBuilder().seed(1).build<double>();
        (void)ica;
    } catch (const DecompositionError& e) {
        std::printf("error: %s\n", e.what());
    }
    return 0;
}
