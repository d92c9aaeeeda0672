// Minimal test macros (GoogleTest is not in the image). Each TEST registers itself; main() runs them all.
#pragma once
#include <cstdio>
#include <functional>
#include <string>
#include <vector>

struct MiniTest {
    static std::vector<std::pair<std::string, std::function<void()>>> &all() {
        static std::vector<std::pair<std::string, std::function<void()>>> v;
        return v;
    }
    static int &failures() { static int f = 0; return f; }
    MiniTest(const char *suite, const char *name, std::function<void()> fn) { all().push_back({std::string(suite) + "." + name, fn}); }
};
#define TEST(suite, name)                                             \
    static void suite##_##name();                                     \
    static MiniTest reg_##suite##_##name(#suite, #name, suite##_##name); \
    static void suite##_##name()
#define EXPECT_TRUE(c) do { if (!(c)) { printf("  FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); ++MiniTest::failures(); } } while (0)
#define EXPECT_EQ(a, b) EXPECT_TRUE((a) == (b))
#define EXPECT_LT(a, b) EXPECT_TRUE((a) < (b))
#define EXPECT_LE(a, b) EXPECT_TRUE((a) <= (b))
#define EXPECT_GE(a, b) EXPECT_TRUE((a) >= (b))
#define MINI_TEST_MAIN()                                                      \
    int main() {                                                              \
        for (auto &t : MiniTest::all()) {                                     \
            int before = MiniTest::failures();                                \
            t.second();                                                       \
            printf("[%s] %s\n", MiniTest::failures() == before ? " OK " : "FAIL", t.first.c_str()); \
        }                                                                     \
        printf("%d failure(s)\n", MiniTest::failures());                      \
        return MiniTest::failures() ? 1 : 0;                                  \
    }
