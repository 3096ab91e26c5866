#pragma once
#include "sais_standin.h"
inline int32_t libsais(const uint8_t* T, int32_t* SA, int32_t n, int32_t /*fs*/, int32_t* /*freq*/) {
    return sais_standin::suffix_sort(T, SA, n);
}
inline int32_t libsais_int(int32_t* T, int32_t* SA, int32_t n, int32_t /*k*/, int32_t /*fs*/) {
    return sais_standin::suffix_sort(T, SA, n);
}
