// CPU check of lidar_processing_v2_b200/csrc/libm_exact.cuh against the running glibc libm.
// Built by tests/test_libm_exact.py with: g++ -O2 -ffp-contract=off -shared -fPIC.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

#include "../../lidar_processing_v2_b200/csrc/libm_exact.cuh"

extern "C"
{
// every float in [lo_bits, hi_bits] (same sign): returns #mismatches of expf
std::uint64_t check_expf_range(std::uint32_t lo_bits, std::uint32_t hi_bits, int threads)
{
    std::vector<std::uint64_t> bad(threads, 0);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t)
    {
        th.emplace_back([&, t]() {
            for (std::uint64_t b = lo_bits + t; b <= hi_bits; b += threads)
            {
                const float x = lpl::u2f(static_cast<std::uint32_t>(b));
                const float a = std::exp(x);
                const float c = lpl::expf_glibc(x);
                if (lpl::f2u(a) != lpl::f2u(c))
                {
                    ++bad[t];
                }
            }
        });
    }
    for (auto& x : th)
    {
        x.join();
    }
    std::uint64_t s = 0;
    for (auto b : bad)
    {
        s += b;
    }
    return s;
}

std::uint64_t check_atanf_range(std::uint32_t lo_bits, std::uint32_t hi_bits, int threads)
{
    std::vector<std::uint64_t> bad(threads, 0);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t)
    {
        th.emplace_back([&, t]() {
            for (std::uint64_t b = lo_bits + t; b <= hi_bits; b += threads)
            {
                const float x = lpl::u2f(static_cast<std::uint32_t>(b));
                if (lpl::f2u(std::atan(x)) != lpl::f2u(lpl::atanf_glibc(x)))
                {
                    ++bad[t];
                }
                const float nx = -x;
                if (lpl::f2u(std::atan(nx)) != lpl::f2u(lpl::atanf_glibc(nx)))
                {
                    ++bad[t];
                }
            }
        });
    }
    for (auto& x : th)
    {
        x.join();
    }
    std::uint64_t s = 0;
    for (auto b : bad)
    {
        s += b;
    }
    return s;
}

// random (y, x) pairs: LiDAR-like magnitudes, millimetre-quantised and raw, plus axis cases
std::uint64_t check_atan2f_random(std::uint64_t count, std::uint32_t seed)
{
    std::mt19937_64 g(seed);
    std::uniform_real_distribution<float> big(-120.f, 120.f);
    std::uniform_real_distribution<float> small(-1e-3f, 1e-3f);
    std::uint64_t bad = 0;
    for (std::uint64_t i = 0; i < count; ++i)
    {
        float y = big(g), x = big(g);
        switch (i & 7)
        {
        case 1:
            y = std::round(y * 1000.f) / 1000.f;
            x = std::round(x * 1000.f) / 1000.f;
            break;
        case 2:
            x = small(g);
            break;
        case 3:
            y = small(g);
            break;
        case 4:
            x = 0.f;
            break;
        case 5:
            y = 0.f;
            break;
        case 6:
            x = 1.0f;
            break;
        default:
            break;
        }
        if (lpl::f2u(std::atan2(y, x)) != lpl::f2u(lpl::atan2f_glibc(y, x)))
        {
            ++bad;
        }
    }
    return bad;
}
} // extern "C"
