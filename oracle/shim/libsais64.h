#pragma once
#include "sais_standin.h"
inline int64_t libsais64(const uint8_t* T, int64_t* SA, int64_t n, int64_t /*fs*/, int64_t* /*freq*/) {
    return sais_standin::suffix_sort(T, SA, n);
}
