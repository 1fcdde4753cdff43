// Host-side row formatter for the streaming result sink: turns
// columns of a result block into the tab-separated rows GWASWriter emits
// (genetest/subscribers.py:192-205), with Python's repr() formatting of
// floats so the file stays byte-compatible with the reference's str(float).
#pragma once

#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstring>

namespace gtk {

// repr(float): shortest digits that round-trip; fixed notation when
// -4 < decpt <= 16, otherwise d.ddde+XX (at least two exponent digits);
// ".0" appended to integral values.
inline char *format_double_repr(double v, char *out) {
    if (std::isnan(v)) { std::memcpy(out, "nan", 3); return out + 3; }
    if (std::isinf(v)) {
        if (v < 0) *out++ = '-';
        std::memcpy(out, "inf", 3);
        return out + 3;
    }
    if (v == 0.0) {
        if (std::signbit(v)) *out++ = '-';
        std::memcpy(out, "0.0", 3);
        return out + 3;
    }
    char tmp[40];
    auto res = std::to_chars(tmp, tmp + sizeof(tmp), v, std::chars_format::scientific);
    // tmp = [-]d[.ddd]e[+-]XX
    char *p = tmp;
    if (*p == '-') { *out++ = '-'; ++p; }
    char digits[24];
    int nd = 0;
    while (p < res.ptr && *p != 'e') {
        if (*p != '.') digits[nd++] = *p;
        ++p;
    }
    ++p;  // 'e'
    int esign = 1;
    if (*p == '-') { esign = -1; ++p; } else if (*p == '+') ++p;
    int ex = 0;
    while (p < res.ptr) ex = ex * 10 + (*p++ - '0');
    ex *= esign;
    const int decpt = ex + 1;          // position of the decimal point
    if (decpt > -4 && decpt <= 16) {
        if (decpt <= 0) {
            *out++ = '0'; *out++ = '.';
            for (int i = 0; i < -decpt; ++i) *out++ = '0';
            for (int i = 0; i < nd; ++i) *out++ = digits[i];
        } else if (decpt >= nd) {
            for (int i = 0; i < nd; ++i) *out++ = digits[i];
            for (int i = nd; i < decpt; ++i) *out++ = '0';
            *out++ = '.'; *out++ = '0';
        } else {
            for (int i = 0; i < decpt; ++i) *out++ = digits[i];
            *out++ = '.';
            for (int i = decpt; i < nd; ++i) *out++ = digits[i];
        }
    } else {
        *out++ = digits[0];
        if (nd > 1) {
            *out++ = '.';
            for (int i = 1; i < nd; ++i) *out++ = digits[i];
        }
        *out++ = 'e';
        *out++ = ex < 0 ? '-' : '+';
        int ae = ex < 0 ? -ex : ex;
        char eb[8];
        int ne = 0;
        while (ae > 0) { eb[ne++] = (char)('0' + ae % 10); ae /= 10; }
        while (ne < 2) eb[ne++] = '0';
        while (ne > 0) *out++ = eb[--ne];
    }
    return out;
}

}  // namespace gtk
