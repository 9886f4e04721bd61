// Minimal stand-in for <doctest/doctest.h> (doctest is not installed in this image): just the four macros the
// reference's C++ test files use -- DOCTEST_CONFIG_IMPLEMENT_WITH_MAIN, TEST_CASE, CHECK, CHECK_THROWS -- so that
// those files compile UNMODIFIED against this repository's fast_pauli.hpp (tests/cpp/Makefile, _ref_tests/).
// Test tooling only; written for this repository, not derived from doctest's sources.
#pragma once
#include <cstdio>
#include <cstring>
#include <exception>
#include <vector>

namespace doctest_shim
{
struct Case
{
    char const *name;
    void (*fn)();
};
inline std::vector<Case> &registry()
{
    static std::vector<Case> r;
    return r;
}
inline int &failures()
{
    static int f = 0;
    return f;
}
inline int &checks()
{
    static int c = 0;
    return c;
}
struct Registrar
{
    Registrar(char const *name, void (*fn)())
    {
        registry().push_back({name, fn});
    }
};
inline void report(bool ok, char const *expr, char const *file, int line)
{
    ++checks();
    if (!ok)
    {
        ++failures();
        std::printf("%s:%d: CHECK( %s ) failed\n", file, line, expr);
    }
}
// run every registered case (optionally only those whose name contains argv[1]); an escaping exception fails the case
inline int run(int argc, char **argv)
{
    int ran = 0;
    for (auto const &c : registry())
    {
        if (argc > 1 && std::strstr(c.name, argv[1]) == nullptr)
            continue;
        ++ran;
        try
        {
            c.fn();
        }
        catch (std::exception const &e)
        {
            ++failures();
            std::printf("TEST_CASE \"%s\" threw: %s\n", c.name, e.what());
        }
    }
    std::printf("[doctest-shim] test cases: %d | assertions: %d | failed: %d\n", ran, checks(), failures());
    return failures() ? 1 : 0;
}
} // namespace doctest_shim

#define DS_CAT2(a, b) a##b
#define DS_CAT(a, b) DS_CAT2(a, b)
#define DS_TEST_CASE_IMPL(fn, name)                                                                                    \
    static void fn();                                                                                                  \
    static doctest_shim::Registrar DS_CAT(fn, _reg)(name, &fn);                                                        \
    static void fn()
#define TEST_CASE(name) DS_TEST_CASE_IMPL(DS_CAT(ds_case_, __COUNTER__), name)
#define CHECK(...) doctest_shim::report(static_cast<bool>(__VA_ARGS__), #__VA_ARGS__, __FILE__, __LINE__)
#define CHECK_THROWS(...)                                                                                              \
    do                                                                                                                 \
    {                                                                                                                  \
        bool ds_threw = false;                                                                                         \
        try                                                                                                            \
        {                                                                                                              \
            static_cast<void>(__VA_ARGS__);                                                                            \
        }                                                                                                              \
        catch (...)                                                                                                    \
        {                                                                                                              \
            ds_threw = true;                                                                                           \
        }                                                                                                              \
        doctest_shim::report(ds_threw, "throws: " #__VA_ARGS__, __FILE__, __LINE__);                                   \
    } while (0)

#ifdef DOCTEST_CONFIG_IMPLEMENT_WITH_MAIN
int main(int argc, char **argv)
{
    return doctest_shim::run(argc, argv);
}
#endif
