// deq_rtc_compat.h — the few libc names the device headers use, so that the same headers compile under nvcc, under
// g++ (host-side unit tests of deq_math.cuh) and under NVRTC (user right-hand sides), which has no system headers.
#pragma once
#if defined(__CUDACC_RTC__)
typedef unsigned char uint8_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#ifndef DBL_EPSILON
#define DBL_EPSILON 2.2204460492503131e-16
#endif
#else
#include <stdint.h>
#include <float.h>
#include <math.h>
#endif
