// main() for a reference test source built against the shim (GoogleTest's gtest_main stand-in).
#include <gtest/gtest.h>

int main(int argc, char** argv) {
    ::testing::InitGoogleTest(&argc, argv);
    return RUN_ALL_TESTS();
}
