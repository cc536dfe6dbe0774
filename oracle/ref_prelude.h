// ref_prelude.h -- force-included (-include) in front of every UNMODIFIED reference translation
// unit when oracle/Makefile compiles /root/reference/src/*.cc in place into oracle/_ref/.
//
// TEST INFRASTRUCTURE ONLY. Nothing in the product (jet-pbrt_b200/) includes this.
//
// It repairs, without touching the reference sources, the four things that stop the reference
// (a Visual-Studio-only project) from compiling with g++:
//   1. missing <memory>/<string> includes (shape.h:57, sampler.h:74, bvh.h:69, film.cc:120);
//   2. FBVH_NodeLeaf<T> is used (bvh.h:69) before it is declared (bvh.h:115);
//   3. PBRT_PRINT("text") expands to log_print("text", ) -- a trailing comma (pbrt.h:30-31).
//      Here log_print becomes a macro that keeps only the format string, so the call is
//      well-formed; the oracle never needs the reference's console output;
//   4. pbrt.cc needs <Windows.h>; it is simply not compiled -- ref_harness.cc supplies
//      log_print_fmt_only() and the app*() timer functions pbrt.h declares.
#pragma once
#include <memory>
#include <string>
#include <limits>
#include <cstdint>

namespace pbrt { template <typename T> class FBVH_NodeLeaf; }

#define JPBRT_REF_FIRST(a, ...) a
#define log_print(...) log_print_fmt_only(JPBRT_REF_FIRST(__VA_ARGS__))
