// gtest_lite -- TEST INFRASTRUCTURE: the handful of GoogleTest macros the reference's test files use (TEST, EXPECT_TRUE/FALSE,
// EXPECT_EQ/NE/LT/LE/GT/GE, ASSERT_*, InitGoogleTest, RUN_ALL_TESTS), so that /root/reference/tests/*.cpp compile unmodified.
// GoogleTest itself is absent from this image. Output format: one "[ RUN/OK/FAILED ] Suite.Name" line per test.
#ifndef GTEST_LITE_H
#define GTEST_LITE_H
#include <cstdio>
#include <functional>
#include <iostream>
#include <string>
#include <vector>

namespace testing {
struct Registry {
    struct Case { std::string name; std::function<void()> fn; };
    static std::vector<Case> &cases() { static std::vector<Case> c; return c; }
    static int &failures() { static int f = 0; return f; }
    static std::string &filter() { static std::string f; return f; }
};
struct Registrar {
    Registrar(const char *suite, const char *name, std::function<void()> fn) {
        Registry::cases().push_back({std::string(suite) + "." + name, fn});
    }
};
inline void InitGoogleTest(int *argc, char **argv) {
    for (int i = 1; i < *argc; ++i) {
        std::string a = argv[i];
        if (a.rfind("--gtest_filter=", 0) == 0) Registry::filter() = a.substr(15);
    }
}
inline int RunAll() {
    int failed = 0, ran = 0;
    for (auto &c : Registry::cases()) {
        if (!Registry::filter().empty() && c.name.find(Registry::filter()) == std::string::npos) continue;
        std::printf("[ RUN      ] %s\n", c.name.c_str());
        std::fflush(stdout);
        const int before = Registry::failures();
        c.fn();
        ++ran;
        if (Registry::failures() != before) { ++failed; std::printf("[  FAILED  ] %s\n", c.name.c_str()); }
        else std::printf("[       OK ] %s\n", c.name.c_str());
        std::fflush(stdout);
    }
    std::printf("[==========] %d tests ran, %d failed\n", ran, failed);
    return failed ? 1 : 0;
}
template <typename A, typename B>
inline void report(const char *file, int line, const char *expr, const A &a, const B &b) {
    ++Registry::failures();
    std::cout << file << ":" << line << ": Failure: " << expr << " with " << a << " vs " << b << std::endl;
}
inline void report1(const char *file, int line, const char *expr) {
    ++Registry::failures();
    std::cout << file << ":" << line << ": Failure: " << expr << std::endl;
}
}  // namespace testing

#define RUN_ALL_TESTS() ::testing::RunAll()
#define TEST(suite, name)                                                                  \
    static void suite##_##name##_body();                                                   \
    static ::testing::Registrar suite##_##name##_reg(#suite, #name, suite##_##name##_body); \
    static void suite##_##name##_body()
#define GTL_CMP(a, b, op, fatal)                                                                     \
    do {                                                                                             \
        const auto gtl_a = (a);                                                                      \
        const auto gtl_b = (b);                                                                      \
        if (!(gtl_a op gtl_b)) {                                                                     \
            ::testing::report(__FILE__, __LINE__, #a " " #op " " #b, gtl_a, gtl_b);                  \
            if (fatal) return;                                                                       \
        }                                                                                            \
    } while (0)
#define EXPECT_TRUE(x) do { if (!(x)) ::testing::report1(__FILE__, __LINE__, "EXPECT_TRUE(" #x ")"); } while (0)
#define EXPECT_FALSE(x) do { if (x) ::testing::report1(__FILE__, __LINE__, "EXPECT_FALSE(" #x ")"); } while (0)
#define ASSERT_TRUE(x) do { if (!(x)) { ::testing::report1(__FILE__, __LINE__, "ASSERT_TRUE(" #x ")"); return; } } while (0)
#define ASSERT_FALSE(x) do { if (x) { ::testing::report1(__FILE__, __LINE__, "ASSERT_FALSE(" #x ")"); return; } } while (0)
#define EXPECT_EQ(a, b) GTL_CMP(a, b, ==, false)
#define EXPECT_NE(a, b) GTL_CMP(a, b, !=, false)
#define EXPECT_LT(a, b) GTL_CMP(a, b, <, false)
#define EXPECT_LE(a, b) GTL_CMP(a, b, <=, false)
#define EXPECT_GT(a, b) GTL_CMP(a, b, >, false)
#define EXPECT_GE(a, b) GTL_CMP(a, b, >=, false)
#define ASSERT_EQ(a, b) GTL_CMP(a, b, ==, true)
#define ASSERT_LT(a, b) GTL_CMP(a, b, <, true)
#define EXPECT_NEAR(a, b, tol) do { const double gtl_d = (double)(a) - (double)(b); if (!(gtl_d <= (tol) && -gtl_d <= (tol))) ::testing::report(__FILE__, __LINE__, "EXPECT_NEAR(" #a ", " #b ")", (a), (b)); } while (0)
#endif
