// capture_formats.hpp -- host side of the C ABI: the file formats the reference's tools use to replay captures
// (SURVEY.md section 8(f) rank 3).  Header only, C++17, no CUDA.
//
//   raw IQ readers   QuantisedIQ<T>::to_c32 + QuantisedIQToFloatIQ<T>::read and get_iq_file_reader_from_mode_string
//                    (vendor/DAB-Radio/examples/app_helpers/app_iq_readers.h:19-87, 107-159): integer samples are
//                    (x - BIAS) * (1 / MAX_AMPLITUDE) with BIAS = 0 / MAX = max() for signed types and
//                    BIAS = MAX = float(max()/2) + 0.5 for unsigned ones; f32 passes through, f64 is narrowed.
//   soft-bit frames  one int8 per bit, as OFDM_Demod emits them (nothing to convert)
//   hard-byte frames convert_viterbi_bits_to_bytes / convert_viterbi_bytes_to_bits
//                    (vendor/DAB-Radio/examples/app_helpers/app_viterbi_convert_block.h:12-44): bit i of byte k is soft bit
//                    8k + i (LSB first); a soft bit >= 0 packs to 1, a 1 unpacks to +127 and a 0 to -127.
// u8 and c32 IQ go to the GPU as they are (the kernels convert u8 on the fly); the other formats are converted here to c32,
// which is what the reference's readers hand to OFDM_Demod::Process.  The "wav" mode of the reference is not provided.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <limits>
#include <type_traits>

namespace dabgpu_host {

enum IqFileFormat { IQF_U8 = 0, IQF_S8, IQF_S16L, IQF_S16B, IQF_U16L, IQF_U16B, IQF_S32L, IQF_S32B, IQF_U32L, IQF_U32B, IQF_F32L, IQF_F32B, IQF_F64L, IQF_F64B,
                    IQF_COUNT };

// the reference's mode strings (app_iq_readers.h:107-113); -1 if unknown or "wav"
inline int iq_format_from_mode(const char* mode) {
    static const char* names[IQF_COUNT] = {"raw_u8", "raw_s8", "raw_s16l", "raw_s16b", "raw_u16l", "raw_u16b", "raw_s32l", "raw_s32b", "raw_u32l", "raw_u32b",
                                           "raw_f32l", "raw_f32b", "raw_f64l", "raw_f64b"};
    for (int i = 0; i < IQF_COUNT; i++)
        if (mode && strcmp(mode, names[i]) == 0) return i;
    return -1;
}

inline size_t iq_component_bytes(int fmt) {
    switch (fmt) {
    case IQF_U8: case IQF_S8: return 1;
    case IQF_S16L: case IQF_S16B: case IQF_U16L: case IQF_U16B: return 2;
    case IQF_F64L: case IQF_F64B: return 8;
    default: return 4;
    }
}

namespace detail {
inline bool machine_little_endian() { const uint16_t v = 1; uint8_t b; memcpy(&b, &v, 1); return b == 1; }
template <typename T> inline T load_component(const uint8_t* p, bool reverse) {
    uint8_t tmp[sizeof(T)];
    for (size_t i = 0; i < sizeof(T); i++) tmp[i] = reverse ? p[sizeof(T) - 1 - i] : p[i];
    T v;
    memcpy(&v, tmp, sizeof(T));
    return v;
}
template <typename T> inline void convert_int(const uint8_t* raw, size_t n_components, bool reverse, float* out) {
    constexpr float BIAS = std::numeric_limits<T>::is_signed ? 0.0f : static_cast<float>(std::numeric_limits<T>::max() / T(2)) + 0.5f;
    constexpr float MAX_AMPLITUDE = std::numeric_limits<T>::is_signed ? static_cast<float>(std::numeric_limits<T>::max())
                                                                       : static_cast<float>(std::numeric_limits<T>::max() / T(2)) + 0.5f;
    constexpr float scale = 1.0f / MAX_AMPLITUDE;
    for (size_t i = 0; i < n_components; i++) {
        const T x = load_component<T>(raw + i * sizeof(T), reverse);
        const float v = (BIAS == 0.0f) ? static_cast<float>(x) : static_cast<float>(x) - BIAS;
        out[i] = v * scale;
    }
}
}  // namespace detail

// raw bytes -> interleaved float I,Q.  Converts whole components only; returns the number of floats written.
inline size_t iq_convert_to_c32(int fmt, const uint8_t* raw, size_t n_bytes, float* out) {
    const size_t cb = iq_component_bytes(fmt);
    const size_t n = n_bytes / cb;
    const bool le = detail::machine_little_endian();
    switch (fmt) {
    case IQF_U8: detail::convert_int<uint8_t>(raw, n, false, out); break;
    case IQF_S8: detail::convert_int<int8_t>(raw, n, false, out); break;
    case IQF_S16L: detail::convert_int<int16_t>(raw, n, !le, out); break;
    case IQF_S16B: detail::convert_int<int16_t>(raw, n, le, out); break;
    case IQF_U16L: detail::convert_int<uint16_t>(raw, n, !le, out); break;
    case IQF_U16B: detail::convert_int<uint16_t>(raw, n, le, out); break;
    case IQF_S32L: detail::convert_int<int32_t>(raw, n, !le, out); break;
    case IQF_S32B: detail::convert_int<int32_t>(raw, n, le, out); break;
    case IQF_U32L: detail::convert_int<uint32_t>(raw, n, !le, out); break;
    case IQF_U32B: detail::convert_int<uint32_t>(raw, n, le, out); break;
    case IQF_F32L: case IQF_F32B:
        for (size_t i = 0; i < n; i++) out[i] = detail::load_component<float>(raw + 4 * i, (fmt == IQF_F32L) != le);
        break;
    case IQF_F64L: case IQF_F64B:
        for (size_t i = 0; i < n; i++) out[i] = static_cast<float>(detail::load_component<double>(raw + 8 * i, (fmt == IQF_F64L) != le));
        break;
    default: return 0;
    }
    return n;
}

// hard-byte frame <-> soft-bit frame (app_viterbi_convert_block.h:12-44)
inline void softbits_to_hard_bytes(const int8_t* bits, size_t n_bytes, uint8_t* bytes) {
    for (size_t k = 0; k < n_bytes; k++) {
        uint8_t v = 0;
        for (int i = 0; i < 8; i++) v |= uint8_t((bits[8 * k + i] >= 0) ? 1 : 0) << i;
        bytes[k] = v;
    }
}
inline void hard_bytes_to_softbits(const uint8_t* bytes, size_t n_bytes, int8_t* bits) {
    for (size_t k = 0; k < n_bytes; k++)
        for (int i = 0; i < 8; i++) bits[8 * k + i] = ((bytes[k] >> i) & 1) ? int8_t(127) : int8_t(-127);
}

}  // namespace dabgpu_host
