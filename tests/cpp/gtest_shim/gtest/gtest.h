// Minimal stand-in for <gtest/gtest.h> (GoogleTest is not in this image): just enough of TEST / TEST_F and the
// EXPECT_* / ASSERT_* macros for the reference's own test sources (cpp/tests/*.cpp of fateshelled/sycl_points) to be
// compiled UNMODIFIED against include/ + libspx.so, so that reference-held assertions run on the CUDA path itself.
// Not a product file: test infrastructure only.
#pragma once
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace testing {
class Test {
public:
    virtual ~Test() {}
    virtual void SetUp() {}
    virtual void TearDown() {}
    virtual void TestBody() = 0;
};
struct Registry {
    struct Entry {
        std::string name;
        std::function<Test*()> make;
    };
    static std::vector<Entry>& all() {
        static std::vector<Entry> v;
        return v;
    }
};
struct Adder {
    Adder(const char* n, std::function<Test*()> f) { Registry::all().push_back({n, f}); }
};
inline int& current_failures() {
    static int f = 0;
    return f;
}
inline bool& current_fatal() {
    static bool f = false;
    return f;
}
inline bool& current_skipped() {
    static bool f = false;
    return f;
}
// streamable failure record; prints at destruction
class Message {
public:
    Message(const char* file, int line, const std::string& what) { s_ << file << ":" << line << ": Failure\n" << what; }
    ~Message() { std::cerr << s_.str() << std::endl; }
    template <typename T>
    Message& operator<<(const T& v) {
        s_ << v;
        return *this;
    }
    Message& operator<<(std::ostream& (*f)(std::ostream&)) {
        s_ << f;
        return *this;
    }

private:
    std::ostringstream s_;
};
struct Void {
    void operator=(const Message&) const {}
};
template <typename T, typename = void>
struct printable : std::false_type {};
template <typename T>
struct printable<T, std::void_t<decltype(std::declval<std::ostream&>() << std::declval<const T&>())>> : std::true_type {};
template <typename T>
std::string show(const T& v) {
    if constexpr (std::is_same_v<T, bool>) {
        return v ? "true" : "false";
    } else if constexpr (printable<T>::value) {
        std::ostringstream s;
        s << v;
        return s.str();
    } else {
        return "<object>";
    }
}
template <typename A, typename B>
std::string cmp_text(const char* op, const char* ea, const char* eb, const A& a, const B& b) {
    return std::string("Expected: (") + ea + ") " + op + " (" + eb + "), actual: " + show(a) + " vs " + show(b) + "\n";
}
inline bool almost_equal_float(float a, float b) {  // 4 ULPs, as EXPECT_FLOAT_EQ
    if (std::isnan(a) || std::isnan(b)) return false;
    if (a == b) return true;
    int32_t ia, ib;
    std::memcpy(&ia, &a, 4);
    std::memcpy(&ib, &b, 4);
    auto biased = [](int32_t i) { return i < 0 ? (uint32_t)(~i + 1) : (uint32_t)i | 0x80000000u; };
    const uint32_t ua = biased(ia), ub = biased(ib);
    return (ua > ub ? ua - ub : ub - ua) <= 4u;
}
inline bool almost_equal_double(double a, double b) {
    if (std::isnan(a) || std::isnan(b)) return false;
    if (a == b) return true;
    int64_t ia, ib;
    std::memcpy(&ia, &a, 8);
    std::memcpy(&ib, &b, 8);
    auto biased = [](int64_t i) { return i < 0 ? (uint64_t)(~i + 1) : (uint64_t)i | 0x8000000000000000ull; };
    const uint64_t ua = biased(ia), ub = biased(ib);
    return (ua > ub ? ua - ub : ub - ua) <= 4u;
}
inline void InitGoogleTest(int*, char**) {}
inline void InitGoogleTest() {}
inline int RunAllTests(const char* filter = nullptr) {
    int failed = 0, passed = 0, skipped = 0;
    for (auto& e : Registry::all()) {
        if (filter && e.name.find(filter) == std::string::npos) continue;
        std::cout << "[ RUN      ] " << e.name << std::endl;
        current_failures() = 0;
        current_fatal() = false;
        current_skipped() = false;
        try {
            Test* t = e.make();
            t->SetUp();
            if (!current_fatal() && !current_skipped()) t->TestBody();
            t->TearDown();
            delete t;
        } catch (const std::exception& ex) {
            std::cerr << "unexpected exception: " << ex.what() << std::endl;
            ++current_failures();
        } catch (...) {
            std::cerr << "unexpected exception" << std::endl;
            ++current_failures();
        }
        if (current_skipped()) {
            ++skipped;
            std::cout << "[  SKIPPED ] " << e.name << std::endl;
        } else if (current_failures()) {
            ++failed;
            std::cout << "[  FAILED  ] " << e.name << std::endl;
        } else {
            ++passed;
            std::cout << "[       OK ] " << e.name << std::endl;
        }
    }
    std::cout << "[==========] " << passed << " passed, " << failed << " failed, " << skipped << " skipped" << std::endl;
    return failed ? 1 : 0;
}
}  // namespace testing

#define RUN_ALL_TESTS() ::testing::RunAllTests()
#define GTEST_CLASS_(a, b) a##_##b##_Test
#define GTEST_DEFINE_(a, b, base)                                                                                    \
    class GTEST_CLASS_(a, b) : public base {                                                                         \
    public:                                                                                                          \
        void TestBody() override;                                                                                    \
    };                                                                                                               \
    static ::testing::Adder a##_##b##_adder(#a "." #b, [] { return (::testing::Test*)new GTEST_CLASS_(a, b)(); });   \
    void GTEST_CLASS_(a, b)::TestBody()
#define TEST(a, b) GTEST_DEFINE_(a, b, ::testing::Test)
#define TEST_F(a, b) GTEST_DEFINE_(a, b, a)

// a failing check records the failure, prints, and (ASSERT) returns from the enclosing void function
#define GTEST_NONFATAL_(what) \
    ++::testing::current_failures(), ::testing::Void() = ::testing::Message(__FILE__, __LINE__, what)
#define GTEST_FATAL_(what) \
    return ++::testing::current_failures(), ::testing::current_fatal() = true, \
           ::testing::Void() = ::testing::Message(__FILE__, __LINE__, what)
#define GTEST_CHECK_(cond, what, fail) \
    if (cond)                          \
        ;                              \
    else                               \
        fail(what)
#define GTEST_CMP_(op, opname, a, b, fail)                                         \
    if (const auto& gt_a_ = (a); true)                                             \
        if (const auto& gt_b_ = (b); true) GTEST_CHECK_(gt_a_ op gt_b_, ::testing::cmp_text(opname, #a, #b, gt_a_, gt_b_), fail)

#define EXPECT_TRUE(c) GTEST_CHECK_((c), std::string("Value of: " #c "\n  Actual: false\nExpected: true\n"), GTEST_NONFATAL_)
#define EXPECT_FALSE(c) GTEST_CHECK_(!(c), std::string("Value of: " #c "\n  Actual: true\nExpected: false\n"), GTEST_NONFATAL_)
#define ASSERT_TRUE(c) GTEST_CHECK_((c), std::string("Value of: " #c "\n  Actual: false\nExpected: true\n"), GTEST_FATAL_)
#define ASSERT_FALSE(c) GTEST_CHECK_(!(c), std::string("Value of: " #c "\n  Actual: true\nExpected: false\n"), GTEST_FATAL_)
#define EXPECT_EQ(a, b) GTEST_CMP_(==, "==", a, b, GTEST_NONFATAL_)
#define EXPECT_NE(a, b) GTEST_CMP_(!=, "!=", a, b, GTEST_NONFATAL_)
#define EXPECT_LT(a, b) GTEST_CMP_(<, "<", a, b, GTEST_NONFATAL_)
#define EXPECT_LE(a, b) GTEST_CMP_(<=, "<=", a, b, GTEST_NONFATAL_)
#define EXPECT_GT(a, b) GTEST_CMP_(>, ">", a, b, GTEST_NONFATAL_)
#define EXPECT_GE(a, b) GTEST_CMP_(>=, ">=", a, b, GTEST_NONFATAL_)
#define ASSERT_EQ(a, b) GTEST_CMP_(==, "==", a, b, GTEST_FATAL_)
#define ASSERT_NE(a, b) GTEST_CMP_(!=, "!=", a, b, GTEST_FATAL_)
#define ASSERT_LT(a, b) GTEST_CMP_(<, "<", a, b, GTEST_FATAL_)
#define ASSERT_LE(a, b) GTEST_CMP_(<=, "<=", a, b, GTEST_FATAL_)
#define ASSERT_GT(a, b) GTEST_CMP_(>, ">", a, b, GTEST_FATAL_)
#define ASSERT_GE(a, b) GTEST_CMP_(>=, ">=", a, b, GTEST_FATAL_)
#define GTEST_NEAR_(a, b, tol, fail)                                                                                   \
    if (const double gt_a_ = (double)(a); true)                                                                        \
        if (const double gt_b_ = (double)(b); true)                                                                    \
            if (const double gt_t_ = (double)(tol); true)                                                              \
    GTEST_CHECK_(std::fabs(gt_a_ - gt_b_) <= gt_t_,                                                                     \
                 std::string("The difference between " #a " and " #b " is ") + std::to_string(std::fabs(gt_a_ - gt_b_)) + \
                     ", which exceeds " #tol " (" + std::to_string(gt_t_) + "); values " + std::to_string(gt_a_) + " vs " + \
                     std::to_string(gt_b_) + "\n",                                                                    \
                 fail)
#define EXPECT_NEAR(a, b, tol) GTEST_NEAR_(a, b, tol, GTEST_NONFATAL_)
#define ASSERT_NEAR(a, b, tol) GTEST_NEAR_(a, b, tol, GTEST_FATAL_)
#define EXPECT_FLOAT_EQ(a, b) \
    GTEST_CHECK_(::testing::almost_equal_float((float)(a), (float)(b)), ::testing::cmp_text("~=", #a, #b, (float)(a), (float)(b)), GTEST_NONFATAL_)
#define ASSERT_FLOAT_EQ(a, b) \
    GTEST_CHECK_(::testing::almost_equal_float((float)(a), (float)(b)), ::testing::cmp_text("~=", #a, #b, (float)(a), (float)(b)), GTEST_FATAL_)
#define EXPECT_DOUBLE_EQ(a, b) \
    GTEST_CHECK_(::testing::almost_equal_double((double)(a), (double)(b)), ::testing::cmp_text("~=", #a, #b, (double)(a), (double)(b)), GTEST_NONFATAL_)
#define ASSERT_DOUBLE_EQ(a, b) \
    GTEST_CHECK_(::testing::almost_equal_double((double)(a), (double)(b)), ::testing::cmp_text("~=", #a, #b, (double)(a), (double)(b)), GTEST_FATAL_)
#define GTEST_THROW_(stmt, extype, fail)                                 \
    if (bool gt_caught_ = false; true) {                                 \
        try {                                                            \
            stmt;                                                        \
        } catch (const extype&) {                                        \
            gt_caught_ = true;                                           \
        } catch (...) {                                                  \
        }                                                                \
        GTEST_CHECK_(gt_caught_, std::string("Expected: " #stmt " throws " #extype "\n"), fail); \
    } else                                                               \
        (void)0
#define EXPECT_THROW(stmt, extype) GTEST_THROW_(stmt, extype, GTEST_NONFATAL_)
#define ASSERT_THROW(stmt, extype) GTEST_THROW_(stmt, extype, GTEST_FATAL_)
#define GTEST_NO_THROW_(stmt, fail)                                      \
    if (bool gt_threw_ = false; true) {                                  \
        try {                                                            \
            stmt;                                                        \
        } catch (...) {                                                  \
            gt_threw_ = true;                                            \
        }                                                                \
        GTEST_CHECK_(!gt_threw_, std::string("Expected: " #stmt " does not throw\n"), fail); \
    } else                                                               \
        (void)0
#define EXPECT_NO_THROW(stmt) GTEST_NO_THROW_(stmt, GTEST_NONFATAL_)
#define ASSERT_NO_THROW(stmt) GTEST_NO_THROW_(stmt, GTEST_FATAL_)
#define FAIL() GTEST_FATAL_(std::string("Failed\n"))
#define ADD_FAILURE() GTEST_NONFATAL_(std::string("Failed\n"))
#define SUCCEED() ::testing::Void() = ::testing::Message(__FILE__, __LINE__, "")
#define GTEST_SKIP() return ::testing::current_skipped() = true, ::testing::Void() = ::testing::Message(__FILE__, __LINE__, "Skipped\n")
