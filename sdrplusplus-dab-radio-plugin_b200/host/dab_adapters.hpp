// dab_adapters.hpp -- C++17 host side above the C ABI (include/dabgpu.h): the reference's own class surfaces for the
// hot path, forwarding to libdabgpu.so with batch = 1.  This is what a maintainer of the plugin links instead of
// ofdm_core / dab_core (INTEGRATION.md shows the CMake change); names, argument meaning and error behaviour follow
// the reference so that code written against it -- and its tests -- read the same.
//
// Reference interfaces mirrored (paths relative to /root/reference):
//   Radio_Block                      src/radio_block.h:11-36, src/radio_block.cpp:11-86
//   OFDM_Demod, OFDM_Demod_Config    vendor/DAB-Radio/src/ofdm/ofdm_demodulator.h:24-141
//   OFDM_Params / get_DAB_OFDM_params vendor/DAB-Radio/src/ofdm/ofdm_params.h, dab_ofdm_params_ref.cpp:10-58
//   DAB_Parameters / get_dab_parameters vendor/DAB-Radio/src/dab/constants/dab_parameters.h:5-90
//   DAB_Viterbi_Decoder              vendor/DAB-Radio/src/dab/algorithms/dab_viterbi_decoder.h:12-45
//   FIC_Decoder                      vendor/DAB-Radio/src/dab/fic/fic_decoder.h:17-38
//   MSC_Decoder, Subchannel          vendor/DAB-Radio/src/dab/msc/msc_decoder.h:19-38, dab/database/dab_database_entities.h
//   Reed_Solomon_Decoder             vendor/DAB-Radio/src/dab/algorithms/reed_solomon_decoder.h:13-27
//   MSC_Reed_Solomon_Data_Packet_Processor vendor/DAB-Radio/src/dab/msc/msc_reed_solomon_data_packet_processor.h:14-45
//   AAC_Frame_Processor              vendor/DAB-Radio/src/dab/audio/aac_frame_processor.h:13-82
//   BasicRadio::Process              vendor/DAB-Radio/src/basic_radio/basic_radio.h:25-56, basic_radio.cpp:41-65
//   Observable                       vendor/DAB-Radio/src/utility/observable.h
//
// Everything lives in namespace dabgpu_host so that a binary may link the reference classes next to these (the
// parity tests do).  Differences that remain are listed in INTEGRATION.md ("what the adapters do not do").
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <complex>
#include <functional>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/dabgpu.h"
#include "fic_autoconfig.hpp"

namespace dabgpu_host {

using viterbi_bit_t = int8_t;

// tcb::span stand-in (pointer + length view)
template <typename T> class span {
    T* m_p = nullptr;
    size_t m_n = 0;
public:
    span() = default;
    span(T* p, size_t n) : m_p(p), m_n(n) {}
    template <typename U, typename = std::enable_if_t<std::is_convertible<U (*)[], T (*)[]>::value>>
    span(const span<U>& o) : m_p(o.data()), m_n(o.size()) {}
    template <typename A, typename U = std::remove_const_t<T>> span(std::vector<U, A>& v) : m_p(v.data()), m_n(v.size()) {}
    template <typename A, typename U = std::remove_const_t<T>, typename = std::enable_if_t<std::is_const<T>::value, U>>
    span(const std::vector<U, A>& v) : m_p(v.data()), m_n(v.size()) {}
    T* data() const { return m_p; }
    size_t size() const { return m_n; }
    bool empty() const { return m_n == 0; }
    T& operator[](size_t i) const { return m_p[i]; }
    T* begin() const { return m_p; }
    T* end() const { return m_p + m_n; }
    span subspan(size_t off, size_t n) const { return span(m_p + off, n); }
    span first(size_t n) const { return span(m_p, n); }
};

// utility/observable.h
template <typename... T> class Observable {
    using Observer = std::function<void(T...)>;
    std::vector<Observer> m_observers;
public:
    void Attach(const Observer& o) { m_observers.push_back(o); }
    void Notify(T... args) { for (const auto& o : m_observers) o(args...); }
    bool empty() const { return m_observers.empty(); }
};

struct OFDM_Params { size_t nb_frame_symbols, nb_symbol_period, nb_null_period, nb_cyclic_prefix, nb_fft, nb_data_carriers; };

struct DAB_Parameters {
    int nb_frame_bits, nb_symbols, nb_fic_symbols, nb_msc_symbols, nb_fibs, nb_cifs, nb_fibs_per_cif;
    int nb_sym_bits, nb_fic_bits, nb_msc_bits, nb_fib_bits, nb_fib_cif_bits, nb_cif_bits;
};

inline dabgpu_params query_params(int transmission_mode) {
    dabgpu_params p;
    if (dabgpu_get_params(transmission_mode, &p) != DABGPU_OK) throw std::runtime_error("Invalid transmission mode");   // same as the reference
    return p;
}

inline OFDM_Params get_DAB_OFDM_params(int transmission_mode) {
    const dabgpu_params p = query_params(transmission_mode);
    return OFDM_Params{size_t(p.nb_frame_symbols), size_t(p.nb_symbol_period), size_t(p.nb_null_period), size_t(p.nb_cyclic_prefix),
                       size_t(p.nb_fft), size_t(p.nb_data_carriers)};
}

inline DAB_Parameters get_dab_parameters(int transmission_mode) {
    const dabgpu_params p = query_params(transmission_mode);
    DAB_Parameters d;
    d.nb_frame_bits = p.nb_frame_bits;
    d.nb_symbols = p.nb_frame_symbols - 1;
    d.nb_sym_bits = d.nb_frame_bits / d.nb_symbols;
    d.nb_fic_symbols = p.nb_fic_bits / d.nb_sym_bits;
    d.nb_msc_symbols = d.nb_symbols - d.nb_fic_symbols;
    d.nb_cifs = p.nb_cifs;
    d.nb_fibs_per_cif = p.nb_fibs_per_cif;
    d.nb_fibs = p.nb_cifs * p.nb_fibs_per_cif;
    d.nb_fic_bits = p.nb_fic_bits;
    d.nb_msc_bits = p.nb_msc_bits;
    d.nb_fib_bits = d.nb_fic_bits / d.nb_fibs;
    d.nb_fib_cif_bits = p.nb_fib_group_bits;
    d.nb_cif_bits = p.nb_cif_bits;
    return d;
}

inline int mode_from_fft(size_t nb_fft) {
    switch (nb_fft) {
    case 2048: return 1;
    case 512: return 2;
    case 256: return 3;
    case 1024: return 4;
    }
    throw std::runtime_error("Invalid transmission mode");
}

// A GPU context shared by the adapter objects of one radio (one stream).  Errors of the C ABI never throw on the hot
// path (reference convention: log and return); they are kept in last_error().
class Context {
    dabgpu_ctx* m_ctx = nullptr;
    std::mutex m_mutex;
    std::string m_last_error;
public:
    Context(int transmission_mode, int iq_format = DABGPU_IQ_C32, unsigned flags = 0, int device = 0, int max_streams = 1) {
        dabgpu_config cfg;
        dabgpu_config_default(&cfg, transmission_mode);
        cfg.device = device;
        cfg.max_streams = max_streams;
        cfg.iq_format = iq_format;
        cfg.flags = flags;
        if (dabgpu_ctx_create(&cfg, &m_ctx) != DABGPU_OK) throw std::runtime_error(std::string("dabgpu_ctx_create: ") + dabgpu_last_error());
    }
    ~Context() { dabgpu_ctx_destroy(m_ctx); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    dabgpu_ctx* get() const { return m_ctx; }
    std::mutex& mutex() { return m_mutex; }
    bool check(int rc) {
        if (rc == DABGPU_OK) return true;
        m_last_error = dabgpu_last_error();
        return false;
    }
    const std::string& last_error() const { return m_last_error; }
};

// Process-wide context for the stateless decoders (Viterbi, FIC group, RS) -- they need a device, not a stream.
inline std::shared_ptr<Context> default_context() {
    static std::mutex m;
    static std::weak_ptr<Context> weak;
    std::lock_guard<std::mutex> lock(m);
    auto sp = weak.lock();
    if (!sp) { sp = std::make_shared<Context>(1, DABGPU_IQ_U8); weak = sp; }
    return sp;
}

// ---------------------------------------------------------------------------------------------------------------
// OFDM_Demod
// ---------------------------------------------------------------------------------------------------------------
struct OFDM_Demod_Config {
    struct { float update_beta = 0.95f; int nb_samples = 100; int nb_decimate = 5; } signal_l1;
    struct { float thresh_null_start = 0.35f; float thresh_null_end = 0.75f; } null_l1_search;
    struct {
        float fine_freq_update_beta = 0.9f;
        bool is_coarse_freq_correction = true;
        float max_coarse_freq_correction_norm = 0.5f;
        float coarse_freq_slow_beta = 0.1f;
        float impulse_peak_threshold_db = 20.0f;
        float impulse_peak_distance_probability = 0.15f;
    } sync;
};

class OFDM_Demod {
public:
    enum State { FINDING_NULL_POWER_DIP, READING_NULL_AND_PRS, RUNNING_COARSE_FREQ_SYNC, RUNNING_FINE_TIME_SYNC, READING_SYMBOLS };
private:
    OFDM_Demod_Config m_cfg, m_cfg_applied;
    const OFDM_Params m_params;
    std::shared_ptr<Context> m_ctx;
    dabgpu_params m_p;
    mutable dabgpu_ofdm_status m_status{};
    std::vector<viterbi_bit_t> m_frame_bits;
    dabgpu_frame_info m_frame_info{};
    // GUI taps: fetched from the device when the getter is called (the render thread polls them, src/render_radio_block.cpp:96-214)
    mutable std::vector<float> m_impulse_response, m_coarse_response;
    mutable std::vector<std::complex<float>> m_frame_fft, m_frame_data_vec, m_correlation_time_buffer;
    Observable<span<const viterbi_bit_t>> m_obs_on_ofdm_frame;
    std::function<void()> m_on_device_frame;   // Radio_Block: channel-decode the frame where it is, on the device
    static bool same(const OFDM_Demod_Config& a, const OFDM_Demod_Config& b) { return memcmp(&a, &b, sizeof(a)) == 0; }
    void push_config() {
        if (same(m_cfg, m_cfg_applied)) return;
        dabgpu_ofdm_config c;
        c.signal_l1_update_beta = m_cfg.signal_l1.update_beta;
        c.signal_l1_nb_samples = m_cfg.signal_l1.nb_samples;
        c.signal_l1_nb_decimate = m_cfg.signal_l1.nb_decimate;
        c.null_thresh_start = m_cfg.null_l1_search.thresh_null_start;
        c.null_thresh_end = m_cfg.null_l1_search.thresh_null_end;
        c.fine_freq_update_beta = m_cfg.sync.fine_freq_update_beta;
        c.is_coarse_freq_correction = m_cfg.sync.is_coarse_freq_correction ? 1 : 0;
        c.max_coarse_freq_correction_norm = m_cfg.sync.max_coarse_freq_correction_norm;
        c.coarse_freq_slow_beta = m_cfg.sync.coarse_freq_slow_beta;
        c.impulse_peak_threshold_db = m_cfg.sync.impulse_peak_threshold_db;
        c.impulse_peak_distance_probability = m_cfg.sync.impulse_peak_distance_probability;
        if (m_ctx->check(dabgpu_ofdm_set_config(m_ctx->get(), &c))) m_cfg_applied = m_cfg;
    }
    void refresh_status() const { dabgpu_ofdm_get_status(m_ctx->get(), 0, &m_status); }
public:
    // prs_fft_ref and carrier_mapper are accepted for signature compatibility; the library builds both tables from the
    // transmission mode (get_DAB_PRS_reference / get_DAB_mapper_ref restated in csrc/ofdm_host.cuh) and the golden
    // tests pin them against the reference's.  nb_desired_threads has no meaning on the GPU.
    OFDM_Demod(const OFDM_Params& params, span<const std::complex<float>> prs_fft_ref, span<const int> carrier_mapper, int nb_desired_threads = 0,
               std::shared_ptr<Context> ctx = nullptr)
        : m_params(params), m_ctx(ctx ? ctx : std::make_shared<Context>(mode_from_fft(params.nb_fft), DABGPU_IQ_C32, DABGPU_FLAG_DIAG_TAPS)) {
        (void)prs_fft_ref; (void)carrier_mapper; (void)nb_desired_threads;
        m_p = query_params(mode_from_fft(params.nb_fft));
        m_frame_bits.resize(size_t(m_p.nb_frame_bits));
        // same extents as the reference's buffers (ofdm_demodulator.cpp:97-110), zero until the first frame
        m_impulse_response.assign(size_t(m_p.nb_fft), 0.0f);
        m_coarse_response.assign(size_t(m_p.nb_fft), 0.0f);
        m_frame_fft.assign(size_t(m_p.nb_frame_symbols + 1) * size_t(m_p.nb_fft), std::complex<float>(0.0f, 0.0f));
        m_frame_data_vec.assign(size_t(m_p.nb_frame_symbols - 1) * size_t(m_p.nb_fft), std::complex<float>(0.0f, 0.0f));
        m_correlation_time_buffer.assign(size_t(m_p.nb_null_period + m_p.nb_symbol_period), std::complex<float>(0.0f, 0.0f));
        memset(&m_cfg_applied, 0, sizeof(m_cfg_applied));
        m_cfg_applied = m_cfg;
    }
    OFDM_Demod(OFDM_Demod&) = delete;
    OFDM_Demod(OFDM_Demod&&) = delete;
    OFDM_Demod& operator=(OFDM_Demod&) = delete;
    OFDM_Demod& operator=(OFDM_Demod&&) = delete;

    // Same contract as the reference: any block length, one producer thread; every call is one Process() block of the
    // sync state machine; observers run before Process returns, with a span that is only valid during the call.
    void Process(span<const std::complex<float>> block) {
        std::lock_guard<std::mutex> lock(m_ctx->mutex());
        push_config();
        if (!m_ctx->check(dabgpu_ofdm_process(m_ctx->get(), block.data(), 0, 0, 1, int(block.size()), int(block.size())))) return;
        if (m_on_device_frame) m_on_device_frame();
        if (m_obs_on_ofdm_frame.empty()) {
            int n = 0;
            dabgpu_ofdm_pop_frames(m_ctx->get(), 0, nullptr, 1 << 20, nullptr, &n);   // keep the queue drained
            return;
        }
        for (;;) {
            int n = 0;
            if (!m_ctx->check(dabgpu_ofdm_pop_frames(m_ctx->get(), 0, m_frame_bits.data(), 1, &m_frame_info, &n)) || n == 0) break;
            m_obs_on_ofdm_frame.Notify(span<const viterbi_bit_t>(m_frame_bits.data(), m_frame_bits.size()));
        }
    }
    void Reset() {
        std::lock_guard<std::mutex> lock(m_ctx->mutex());
        m_ctx->check(dabgpu_ofdm_reset(m_ctx->get(), 0));
    }
    OFDM_Params GetOFDMParams() const { return m_params; }
    State GetState() const { refresh_status(); return State(m_status.state); }
    auto& GetConfig() { return m_cfg; }
    const auto& GetConfig() const { return m_cfg; }
    float GetSignalAverage() const { refresh_status(); return m_status.signal_l1_average; }
    float GetFineFrequencyOffset() const { refresh_status(); return m_status.freq_fine_offset; }
    float GetCoarseFrequencyOffset() const { refresh_status(); return m_status.freq_coarse_offset; }
    float GetNetFrequencyOffset() const { refresh_status(); return m_status.freq_fine_offset + m_status.freq_coarse_offset; }
    int GetFineTimeOffset() const { refresh_status(); return m_status.fine_time_offset; }
    int GetTotalFramesRead() const { refresh_status(); return m_status.total_frames_read; }
    int GetTotalFramesDesync() const { refresh_status(); return m_status.total_frames_desync; }
    span<const viterbi_bit_t> GetFrameDataBits() const { return span<const viterbi_bit_t>(m_frame_bits.data(), m_frame_bits.size()); }
    // The views below are refreshed from the device by the call and stay valid until the next call of the same getter.  They keep
    // their previous content when the context has no taps (a caller-supplied Context without DABGPU_FLAG_DIAG_TAPS) or no frame
    // was emitted yet.
    span<const std::complex<float>> GetFrameFFT() const {   // (nb_frame_symbols + 1) x nb_fft, NULL symbol last
        std::lock_guard<std::mutex> lock(m_ctx->mutex());
        dabgpu_ofdm_get_frame_fft(m_ctx->get(), 0, reinterpret_cast<float*>(m_frame_fft.data()), m_frame_fft.size() * 2);
        return span<const std::complex<float>>(m_frame_fft.data(), m_frame_fft.size());
    }
    span<const std::complex<float>> GetFrameDataVec() const {   // the reference's extent (L-1) x nb_fft, (L-1) x nb_data_carriers used
        std::lock_guard<std::mutex> lock(m_ctx->mutex());
        dabgpu_ofdm_get_frame_data_vec(m_ctx->get(), 0, reinterpret_cast<float*>(m_frame_data_vec.data()), m_frame_data_vec.size() * 2);
        return span<const std::complex<float>>(m_frame_data_vec.data(), m_frame_data_vec.size());
    }
    span<const float> GetImpulseResponse() const {
        std::lock_guard<std::mutex> lock(m_ctx->mutex());
        dabgpu_ofdm_get_response(m_ctx->get(), 0, 0, m_impulse_response.data(), int(m_impulse_response.size()));
        return span<const float>(m_impulse_response.data(), m_impulse_response.size());
    }
    span<const float> GetCoarseFrequencyResponse() const {
        std::lock_guard<std::mutex> lock(m_ctx->mutex());
        dabgpu_ofdm_get_response(m_ctx->get(), 0, 1, m_coarse_response.data(), int(m_coarse_response.size()));
        return span<const float>(m_coarse_response.data(), m_coarse_response.size());
    }
    span<const std::complex<float>> GetCorrelationTimeBuffer() const {
        std::lock_guard<std::mutex> lock(m_ctx->mutex());
        dabgpu_ofdm_get_correlation_buffer(m_ctx->get(), 0, reinterpret_cast<float*>(m_correlation_time_buffer.data()), m_correlation_time_buffer.size() * 2);
        return span<const std::complex<float>>(m_correlation_time_buffer.data(), m_correlation_time_buffer.size());
    }
    auto& On_OFDM_Frame() { return m_obs_on_ofdm_frame; }
    // not in the reference: offsets in force when the frame being delivered to the observers was demodulated
    const dabgpu_frame_info& GetFrameInfo() const { return m_frame_info; }
    // not in the reference: used by Radio_Block to keep the soft bits on the device
    std::shared_ptr<Context> GetContext() const { return m_ctx; }
    void SetDeviceFrameHook(std::function<void()> f) { m_on_device_frame = std::move(f); }
};

// ofdm/ofdm_helpers.h:12-20
inline std::unique_ptr<OFDM_Demod> Create_OFDM_Demodulator(int transmission_mode, int total_threads = 0) {
    const OFDM_Params p = get_DAB_OFDM_params(transmission_mode);
    return std::make_unique<OFDM_Demod>(p, span<const std::complex<float>>(), span<const int>(), total_threads);
}

// ---------------------------------------------------------------------------------------------------------------
// DAB_Viterbi_Decoder: reset(); update(...)*; chainback()  ==  one dabgpu_viterbi_job
// ---------------------------------------------------------------------------------------------------------------
class DAB_Viterbi_Decoder {
public:
    static constexpr size_t m_constraint_length = 7;
    static constexpr size_t m_code_rate = 4;
private:
    std::shared_ptr<Context> m_ctx;
    std::vector<int8_t> m_soft;
    dabgpu_viterbi_job m_job{};
    size_t m_traceback_length = 0;
    size_t m_decoded_bits = 0;
    // EN 300 401 table 13 in the reference's count form (puncture_codes.h:42-69): PI_i keeps 8+i of 32 mother bits
    static void pi_pattern(int pi, uint8_t out[32]) {
        static const int order[8] = {0, 4, 2, 6, 1, 5, 3, 7};
        int cnt[8];
        const int base = 1 + (pi - 1) / 8;
        for (int g = 0; g < 8; g++) cnt[g] = base;
        for (int k = 0; k <= (pi - 1) % 8; k++) cnt[order[k]]++;
        for (int g = 0; g < 8; g++)
            for (int r = 0; r < 4; r++) out[4 * g + r] = r < cnt[g] ? 1 : 0;
    }
    // 1..24 = PI table row, 0 = tail code, -1 = a pattern the tables do not contain
    static int identify(span<const uint8_t> code) {
        if (code.size() == 24) {
            bool tail = true;
            for (size_t i = 0; i < 24; i++) tail = tail && (code[i] != 0) == ((i & 3) < 2);
            if (tail) return 0;
        }
        if (code.size() == 32) {
            for (int pi = 1; pi <= 24; pi++) {
                uint8_t pat[32];
                pi_pattern(pi, pat);
                bool eq = true;
                for (int i = 0; i < 32; i++) eq = eq && ((code[size_t(i)] != 0) == (pat[i] != 0));
                if (eq) return pi;
            }
        }
        return -1;
    }
public:
    explicit DAB_Viterbi_Decoder(std::shared_ptr<Context> ctx = nullptr) : m_ctx(ctx ? ctx : default_context()) { reset(); }
    void set_traceback_length(const size_t traceback_length) { m_traceback_length = traceback_length; }
    size_t get_traceback_length() const { return m_traceback_length; }
    size_t get_current_decoded_bit() const { return m_decoded_bits; }
    void reset(const size_t starting_state = 0u) {
        (void)starting_state;   // the DAB chain always starts in state 0 (fic_decoder.cpp:74, msc_decoder.cpp:84)
        m_soft.clear();
        memset(&m_job, 0, sizeof(m_job));
        m_decoded_bits = 0;
    }
    // returns the number of punctured symbols consumed, like the reference (dab_viterbi_decoder.cpp:109-123)
    size_t update(span<const viterbi_bit_t> punctured_symbols, span<const uint8_t> puncture_code, const size_t requested_output_symbols) {
        if (requested_output_symbols == 0 || puncture_code.empty()) return 0;
        size_t consumed = 0;
        for (size_t i = 0; i < requested_output_symbols; i++) consumed += puncture_code[i % puncture_code.size()] ? 1 : 0;
        if (consumed > punctured_symbols.size() || m_job.n_seg >= DABGPU_MAX_SEGMENTS || requested_output_symbols % m_code_rate != 0) return 0;
        const int pi = identify(puncture_code);
        const uint32_t k = m_job.n_seg++;
        if (pi >= 0) {
            m_job.seg_pi[k] = uint8_t(pi);
            m_soft.insert(m_soft.end(), punctured_symbols.data(), punctured_symbols.data() + consumed);
        } else {
            // a code outside table 13: hand the symbols over already de-punctured, as the unpunctured code PI_24
            m_job.seg_pi[k] = 24;
            size_t j = 0;
            for (size_t i = 0; i < requested_output_symbols; i++) m_soft.push_back(puncture_code[i % puncture_code.size()] ? punctured_symbols[j++] : int8_t(0));
        }
        m_job.seg_bits[k] = uint32_t(requested_output_symbols);
        m_decoded_bits += requested_output_symbols / m_code_rate;
        return consumed;
    }
    // returns the accumulated path error, like the reference (dab_viterbi_decoder.cpp:125-129)
    uint64_t chainback(span<uint8_t> bytes_out, const size_t end_state = 0u) {
        (void)end_state;   // always 0 in the DAB chain
        m_job.soft_offset = 0;
        m_job.n_soft = uint32_t(m_soft.size());
        m_job.out_offset = 0;
        m_job.n_out_bytes = uint32_t(bytes_out.size());
        m_job.descramble = 0;
        uint64_t err = 0;
        std::lock_guard<std::mutex> lock(m_ctx->mutex());
        m_ctx->check(dabgpu_viterbi_decode(m_ctx->get(), &m_job, 1, m_soft.data(), m_soft.size(), bytes_out.data(), bytes_out.size(), &err));
        return err;
    }
};

// ---------------------------------------------------------------------------------------------------------------
// FIC_Decoder
// ---------------------------------------------------------------------------------------------------------------
class FIC_Decoder {
    std::shared_ptr<Context> m_ctx;
    std::vector<uint8_t> m_decoded_bytes;
    const size_t m_nb_fibs_per_group, m_nb_encoded_bits;
    Observable<span<const uint8_t>> obs_on_fib;
public:
    FIC_Decoder(const size_t nb_encoded_bits, const size_t nb_fibs_per_group, std::shared_ptr<Context> ctx = nullptr)
        : m_ctx(ctx ? ctx : default_context()), m_decoded_bytes(96), m_nb_fibs_per_group(nb_fibs_per_group), m_nb_encoded_bits(nb_encoded_bits) {}
    void DecodeFIBGroup(span<const viterbi_bit_t> encoded_bits, const size_t cif_index) {
        (void)cif_index;
        // like the reference, only 2304-bit groups of 3 FIBs are decodable (fic_decoder.cpp:66-72); others are dropped
        if (m_nb_encoded_bits != 2304 || encoded_bits.size() != 2304 || m_nb_fibs_per_group != 3) return;
        uint8_t ok[3] = {0, 0, 0};
        {
            std::lock_guard<std::mutex> lock(m_ctx->mutex());
            if (!m_ctx->check(dabgpu_fic_decode(m_ctx->get(), encoded_bits.data(), 1, m_decoded_bytes.data(), ok))) return;
        }
        for (size_t i = 0; i < 3; i++)
            if (ok[i]) obs_on_fib.Notify(span<const uint8_t>(m_decoded_bytes.data() + 32 * i, 30));   // CRC failures are not forwarded
    }
    auto& OnFIB(void) { return obs_on_fib; }
};

// ---------------------------------------------------------------------------------------------------------------
// MSC_Decoder
// ---------------------------------------------------------------------------------------------------------------
enum class EEP_Type : uint8_t { TYPE_A, TYPE_B };
typedef uint8_t subchannel_id_t;
struct Subchannel {   // dab/database/dab_database_entities.h (fields used by the decoder)
    subchannel_id_t id = 0;
    uint16_t start_address = 0;
    uint16_t length = 0;
    bool is_uep = false;
    uint8_t uep_prot_index = 0;
    uint8_t eep_prot_level = 0;
    EEP_Type eep_type = EEP_Type::TYPE_A;
};

inline dabgpu_subchannel to_abi(const Subchannel& s, bool is_dabplus) {
    dabgpu_subchannel d;
    d.start_address = s.start_address;
    d.length = s.length;
    d.is_uep = s.is_uep ? 1 : 0;
    d.uep_prot_index = s.uep_prot_index;
    d.eep_prot_level = s.eep_prot_level;
    d.eep_type_b = s.eep_type == EEP_Type::TYPE_B ? 1 : 0;
    d.is_dabplus = is_dabplus ? 1 : 0;
    return d;
}

// The stand-alone MSC_Decoder objects of a process share GPU contexts: a pool context has the single-CIF-per-frame geometry
// (transmission mode II, FIC disabled) that gives DecodeCIF its one-CIF-at-a-time contract, and MSC_POOL_STREAMS streams; every
// MSC_Decoder leases one stream of it (its own frame ring = its own CIF_Deinterleaver history).  A full pool is followed by
// another one.  The reference builds one MSC_Decoder per sub-channel of a radio (basic_audio_channel.cpp), i.e. tens per process.
constexpr int MSC_POOL_STREAMS = 32;
class MscPool {
    std::shared_ptr<Context> m_ctx;
    std::vector<bool> m_used;
public:
    MscPool() : m_ctx(std::make_shared<Context>(2, DABGPU_IQ_U8, DABGPU_FLAG_NO_FIC, 0, MSC_POOL_STREAMS)), m_used(size_t(MSC_POOL_STREAMS), false) {}
    const std::shared_ptr<Context>& context() const { return m_ctx; }
    struct Lease { std::shared_ptr<MscPool> pool; int stream = -1; };
    static Lease acquire() {
        static std::mutex m;
        static std::vector<std::weak_ptr<MscPool>> pools;
        std::lock_guard<std::mutex> lock(m);
        for (auto& w : pools)
            if (auto p = w.lock())
                for (int i = 0; i < MSC_POOL_STREAMS; i++)
                    if (!p->m_used[size_t(i)]) { p->m_used[size_t(i)] = true; return Lease{p, i}; }
        auto p = std::make_shared<MscPool>();
        pools.push_back(p);
        p->m_used[0] = true;
        return Lease{p, 0};
    }
    void release(int stream) {
        std::lock_guard<std::mutex> lock(m_ctx->mutex());
        dabgpu_msc_configure(m_ctx->get(), stream, nullptr, 0);
        m_used[size_t(stream)] = false;   // a later lease reconfigures the stream; its de-interleaver restarts empty
    }
};

// One sub-channel of the common interleaved frame: CIF_Deinterleaver + EEP/UEP de-puncture + Viterbi + descramble.
// The time de-interleaver history lives in the soft-bit frame ring of the leased stream.
class MSC_Decoder {
    const Subchannel m_subchannel;
    MscPool::Lease m_lease;
    std::shared_ptr<Context> m_ctx;
    std::vector<viterbi_bit_t> m_frame;
    std::vector<uint8_t> m_decoded_bytes_buf;
    dabgpu_params m_p;
    bool m_ok = false;
public:
    explicit MSC_Decoder(const Subchannel subchannel) : m_subchannel(subchannel), m_lease(MscPool::acquire()), m_ctx(m_lease.pool->context()) {
        m_p = query_params(2);
        m_frame.assign(size_t(m_p.nb_frame_bits), 0);
        const dabgpu_subchannel d = to_abi(subchannel, false);
        std::lock_guard<std::mutex> lock(m_ctx->mutex());
        m_ok = m_ctx->check(dabgpu_msc_configure(m_ctx->get(), m_lease.stream, &d, 1));
        int nb = 0;
        if (m_ok) dabgpu_msc_get_layout(m_ctx->get(), m_lease.stream, 0, nullptr, &nb);
        m_decoded_bytes_buf.resize(size_t(nb));
    }
    ~MSC_Decoder() { m_lease.pool->release(m_lease.stream); }
    MSC_Decoder(const MSC_Decoder&) = delete;
    MSC_Decoder& operator=(const MSC_Decoder&) = delete;
    // Returns a view of the decoded bytes; empty while the 16-CIF de-interleaver is still filling (msc_decoder.cpp:46-75)
    span<uint8_t> DecodeCIF(span<const viterbi_bit_t> buf) {
        if (!m_ok || buf.size() != size_t(m_p.nb_cif_bits)) return span<uint8_t>();
        memcpy(m_frame.data() + m_p.nb_fic_bits, buf.data(), buf.size());
        const int s = m_lease.stream;
        std::lock_guard<std::mutex> lock(m_ctx->mutex());
        if (!m_ctx->check(dabgpu_softbits_push(m_ctx->get(), m_frame.data(), m_frame.size(), s, 1))) return span<uint8_t>();
        if (!m_ctx->check(dabgpu_chan_decode(m_ctx->get(), s, 1))) return span<uint8_t>();
        uint8_t valid = 0;
        int nb = 0;
        if (!m_ctx->check(dabgpu_chan_get_msc(m_ctx->get(), s, 0, m_decoded_bytes_buf.data(), m_decoded_bytes_buf.size(), &valid, &nb))) return span<uint8_t>();
        return valid ? span<uint8_t>(m_decoded_bytes_buf.data(), size_t(nb)) : span<uint8_t>();
    }
    bool IsValid() const { return m_ok; }
    const std::string& LastError() const { return m_ctx->last_error(); }
};

// ---------------------------------------------------------------------------------------------------------------
// Reed_Solomon_Decoder  (Phil Karn's decode_rs_char semantics: returns #corrected symbols or -1; positions incl. pad)
// ---------------------------------------------------------------------------------------------------------------
class Reed_Solomon_Decoder {
    std::shared_ptr<Context> m_ctx;
    int m_nroots, m_pad;
    bool m_supported;
public:
    Reed_Solomon_Decoder(const int symbol_size, const int galois_field_polynomial, const int fcr, const int primer, const int nb_roots, const int pad,
                         std::shared_ptr<Context> ctx = nullptr)
        : m_ctx(ctx ? ctx : default_context()), m_nroots(nb_roots), m_pad(pad),
          // the two codes of the DAB chain: GF(2^8)/0x11D, fcr 0, prim 1 (aac_frame_processor.cpp:105-111, msc_reed_solomon_data_packet_processor.cpp:21-26)
          m_supported(symbol_size == 8 && galois_field_polynomial == 0x11D && fcr == 0 && primer == 1 && nb_roots >= 1 && nb_roots <= 32 && pad >= 0 &&
                      pad < 255 - nb_roots) {}
    Reed_Solomon_Decoder(Reed_Solomon_Decoder&) = delete;
    Reed_Solomon_Decoder& operator=(Reed_Solomon_Decoder&) = delete;
    int Decode(uint8_t* data, int* eras_pos, int no_eras) {
        if (!m_supported || no_eras != 0) return -1;   // the DAB chain never passes erasures (aac_frame_processor.cpp:343)
        int count = -1;
        std::vector<int> pos(size_t(m_nroots), 0);
        std::lock_guard<std::mutex> lock(m_ctx->mutex());
        if (!m_ctx->check(dabgpu_rs_decode(m_ctx->get(), data, 1, m_nroots, m_pad, &count, pos.data()))) return -1;
        if (eras_pos) for (int i = 0; i < count; i++) eras_pos[i] = pos[size_t(i)];
        return count;
    }
};

// ---------------------------------------------------------------------------------------------------------------
// MSC_Reed_Solomon_Data_Packet_Processor: packet-mode FEC (ETSI EN 300 401 5.3.5).  Packets are queued until the nine FEC
// packets of a set have arrived; the Reed-Solomon step over the 12 rows of the set runs on the GPU
// (dabgpu_packet_fec_decode), the packets then leave through the callback flagged "corrected".  Packets that cannot be
// protected (a broken FEC counter sequence, an incomplete set) leave flagged "not corrected", in arrival order, like the
// reference (msc_reed_solomon_data_packet_processor.cpp:49-258).
// ---------------------------------------------------------------------------------------------------------------
class MSC_Reed_Solomon_Data_Packet_Processor {
public:
    using Callback = std::function<void(span<const uint8_t>, bool)>;   // packet, is_corrected
private:
    static constexpr size_t kTableBytes = 2256, kFecPackets = 9, kFecPacketBytes = 24, kFecHeader = 2;
    static constexpr size_t kCapacity = kTableBytes + kFecPackets * kFecPacketBytes;
    static constexpr uint16_t kFecAddress = 0x3FE;
    static size_t length_of(uint8_t id) { return 24u * (size_t(id & 3u) + 1u); }   // table 6: 24, 48, 72, 96

    std::shared_ptr<Context> m_ctx;
    std::vector<uint8_t> m_store = std::vector<uint8_t>(kCapacity);   // circular queue of whole packets
    size_t m_front = 0, m_back = 0, m_fill = 0;
    int m_fec_seen = -1;                                              // counter of the last FEC packet of the running set
    std::vector<uint8_t> m_packet, m_frame = std::vector<uint8_t>(DABGPU_PACKET_FEC_FRAME_BYTES);
    Callback m_callback;

    uint8_t& cell(size_t offset_from_front) { return m_store[(m_front + offset_from_front) % kCapacity]; }
    void enqueue(span<const uint8_t> packet, uint8_t length_id) {
        while (kCapacity - m_fill < packet.size()) {   // the oldest packets make room, unannounced
            const size_t n = length_of(uint8_t(m_store[m_front] >> 6));
            m_fill -= n;
            m_front = (m_front + n) % kCapacity;
        }
        for (size_t i = 0; i < packet.size(); i++) m_store[(m_back + i) % kCapacity] = packet[i];
        m_store[m_back] = uint8_t((packet[0] & 0x3Fu) | (length_id << 6));   // a FEC packet is always stored as 24 bytes
        m_fill += packet.size();
        m_back = (m_back + packet.size()) % kCapacity;
    }
    bool dequeue() {
        if (m_fill == 0) return false;
        const size_t n = length_of(uint8_t(m_store[m_front] >> 6));
        m_packet.resize(n);
        for (size_t i = 0; i < n; i++) m_packet[i] = cell(i);
        m_fill -= n;
        m_front = (m_front + n) % kCapacity;
        return true;
    }
    void flush_uncorrected() {
        while (dequeue())
            if (m_callback) m_callback({m_packet.data(), m_packet.size()}, false);
    }
    void correct_and_flush() {
        // FEC frame for the device: the application data table as queued, then the data fields of the nine FEC packets
        for (size_t i = 0; i < kTableBytes; i++) m_frame[i] = cell(i);
        size_t w = kTableBytes;
        for (size_t k = 0; k < kFecPackets; k++) {
            const size_t n = (k + 1 < kFecPackets) ? kFecPacketBytes - kFecHeader : kFecPacketBytes - kFecHeader - 6;   // six padding bytes end the set
            for (size_t j = 0; j < n; j++) m_frame[w++] = cell(kTableBytes + k * kFecPacketBytes + kFecHeader + j);
        }
        bool ok;
        {
            std::lock_guard<std::mutex> lock(m_ctx->mutex());
            ok = m_ctx->check(dabgpu_packet_fec_decode(m_ctx->get(), m_frame.data(), 1, nullptr));
        }
        if (ok) for (size_t i = 0; i < kTableBytes; i++) cell(i) = m_frame[i];
        size_t delivered = 0;
        while (delivered < kTableBytes && dequeue()) {
            if (m_callback) m_callback({m_packet.data(), m_packet.size()}, true);
            delivered += m_packet.size();
        }
    }
public:
    explicit MSC_Reed_Solomon_Data_Packet_Processor(std::shared_ptr<Context> ctx = nullptr) : m_ctx(ctx ? ctx : default_context()) {}
    void SetCallback(const Callback& callback) { m_callback = callback; }
    void SetCallback(Callback&& callback) { m_callback = std::move(callback); }
    size_t ReadPacket(span<const uint8_t> buf) {
        if (buf.size() < kFecHeader) return buf.size();
        const uint16_t address = uint16_t((uint16_t(buf[0] & 3u) << 8) | buf[1]);
        const uint8_t counter = uint8_t((buf[0] >> 2) & 0xFu);
        const bool is_fec = address == kFecAddress;
        const uint8_t length_id = is_fec ? uint8_t(0) : uint8_t(buf[0] >> 6);   // the length field of a FEC packet is not trusted
        const size_t n = length_of(length_id);
        if (buf.size() < n) return buf.size();
        enqueue(buf.first(n), length_id);
        if (!is_fec) return n;
        if (int(counter) != m_fec_seen + 1) {      // sets count 0..8; anything else abandons the set
            m_fec_seen = -1;
            flush_uncorrected();
            return n;
        }
        m_fec_seen = int(counter);
        if (size_t(counter) + 1 != kFecPackets) return n;
        if (m_fill == kCapacity) correct_and_flush(); else flush_uncorrected();
        m_fec_seen = -1;
        m_front = m_back = m_fill = 0;
        return n;
    }
};

// ---------------------------------------------------------------------------------------------------------------
// AAC_Frame_Processor
// ---------------------------------------------------------------------------------------------------------------
enum class MPEG_Surround { NOT_USED, SURROUND_51, SURROUND_71, SURROUND_OTHER, RFA };
struct SuperFrameHeader {
    uint32_t sampling_rate = 0;
    bool is_parametric_stereo = false;
    bool is_spectral_band_replication = false;
    bool is_stereo = false;
    MPEG_Surround mpeg_surround = MPEG_Surround::NOT_USED;
    bool operator==(const SuperFrameHeader& o) const {
        return sampling_rate == o.sampling_rate && is_parametric_stereo == o.is_parametric_stereo &&
               is_spectral_band_replication == o.is_spectral_band_replication && is_stereo == o.is_stereo && mpeg_surround == o.mpeg_surround;
    }
    bool operator!=(const SuperFrameHeader& o) const { return !(*this == o); }
};

// Fires the observers of a flat event log (dabgpu_event_header records) in order
struct DabPlusObservers {
    Observable<const int, const uint16_t, const uint16_t> firecode_error;
    Observable<const int, const int> rs_error;
    Observable<SuperFrameHeader> superframe_header;
    Observable<const int, const int, const uint16_t, const uint16_t> au_crc_error;
    Observable<const int, const int, span<uint8_t>> access_unit;
    void dispatch(uint8_t* log, size_t n) {
        size_t off = 0;
        while (off + sizeof(dabgpu_event_header) <= n) {
            dabgpu_event_header h;
            memcpy(&h, log + off, sizeof(h));
            off += sizeof(h);
            uint8_t* payload = log + off;
            off += (size_t(h.n_bytes) + 3u) & ~size_t(3);
            switch (h.type) {
            case DABGPU_EV_FIRECODE_ERROR: firecode_error.Notify(h.a, uint16_t(h.b), uint16_t(h.c)); break;
            case DABGPU_EV_RS_ERROR: rs_error.Notify(h.a, h.b); break;
            case DABGPU_EV_SUPERFRAME_HEADER: {
                SuperFrameHeader sh;
                sh.sampling_rate = uint32_t(h.a);
                sh.is_parametric_stereo = (h.b & 1) != 0;
                sh.is_spectral_band_replication = (h.b & 2) != 0;
                sh.is_stereo = (h.b & 4) != 0;
                sh.mpeg_surround = MPEG_Surround(h.c);
                superframe_header.Notify(sh);
                break;
            }
            case DABGPU_EV_AU_CRC_ERROR: au_crc_error.Notify(h.a, h.b, uint16_t(h.c), uint16_t(h.d)); break;
            case DABGPU_EV_ACCESS_UNIT: access_unit.Notify(h.a, h.b, span<uint8_t>(payload, size_t(h.n_bytes))); break;
            default: break;
            }
        }
    }
};

class AAC_Frame_Processor {
    std::shared_ptr<Context> m_ctx;
    dabgpu_dabplus* m_proc = nullptr;
    std::vector<uint8_t> m_log;
    DabPlusObservers m_obs;
public:
    explicit AAC_Frame_Processor(std::shared_ptr<Context> ctx = nullptr) : m_ctx(ctx ? ctx : default_context()), m_log(1 << 16) {
        if (dabgpu_dabplus_open(m_ctx->get(), &m_proc) != DABGPU_OK) throw std::runtime_error(std::string("dabgpu_dabplus_open: ") + dabgpu_last_error());
    }
    ~AAC_Frame_Processor() { dabgpu_dabplus_close(m_ctx->get(), m_proc); }
    AAC_Frame_Processor(const AAC_Frame_Processor&) = delete;
    AAC_Frame_Processor& operator=(const AAC_Frame_Processor&) = delete;
    // An audio super frame consists of 5 DAB logical frames (aac_frame_processor.cpp:126-177)
    void Process(span<const uint8_t> buf) {
        size_t n = 0;
        {
            std::lock_guard<std::mutex> lock(m_ctx->mutex());
            if (!m_ctx->check(dabgpu_dabplus_process(m_ctx->get(), m_proc, buf.data(), int(buf.size()), m_log.data(), m_log.size(), &n))) return;
        }
        m_obs.dispatch(m_log.data(), n);
    }
    auto& OnFirecodeError(void) { return m_obs.firecode_error; }
    auto& OnRSError(void) { return m_obs.rs_error; }
    auto& OnSuperFrameHeader(void) { return m_obs.superframe_header; }
    auto& OnAccessUnitCRCError(void) { return m_obs.au_crc_error; }
    auto& OnAccessUnit(void) { return m_obs.access_unit; }
};

// ---------------------------------------------------------------------------------------------------------------
// BasicRadio: the frame-level seam (BasicRadio::Process, basic_radio.cpp:41-65).  FIG parsing / the DAB database
// that make the reference self-configuring are outside the hot path: sub-channels are
// declared with SetSubchannels, everything they decode is delivered through observers.
// ---------------------------------------------------------------------------------------------------------------
class BasicRadio {
public:
    struct Channel {
        Subchannel subchannel;
        bool is_dabplus = true;
        Observable<span<const uint8_t>> on_msc_data;   // one call per decoded logical frame (MSC_Decoder::DecodeCIF output)
        DabPlusObservers dabplus;                       // AAC_Frame_Processor observers of that sub-channel
    };
private:
    const DAB_Parameters m_params;
    std::shared_ptr<Context> m_ctx;
    std::mutex m_mutex_data;
    std::vector<std::unique_ptr<Channel>> m_channels;
    Observable<span<const uint8_t>> m_obs_on_fib;
    Observable<subchannel_id_t, Channel&> m_obs_audio_channel;
    std::vector<uint8_t> m_fibs, m_crc, m_msc, m_log;
    FIC_Autoconfig m_autocfg;
    bool m_self_configure = false;
    std::vector<uint8_t> m_configured_ids;
    std::vector<uint8_t> m_rejected_ids;
    // BasicRadio::UpdateAfterProcessing (basic_radio.cpp:83-154): every newly complete audio sub-channel gets a decoder
    // (dabgpu_msc_add_subchannel) and the observers are told; the decoders already running keep their de-interleaver and
    // superframe state.  A sub-channel the context refuses only affects itself and is not retried.
    void update_after_processing() {
        std::vector<dabgpu_subchannel> subs;
        std::vector<uint8_t> ids;
        m_autocfg.Runnable(subs, ids);
        for (size_t i = 0; i < subs.size(); i++) {
            bool known = false;
            for (uint8_t k : m_configured_ids) known = known || (k == ids[i]);
            for (uint8_t k : m_rejected_ids) known = known || (k == ids[i]);
            if (known) continue;
            int index = -1;
            if (!m_ctx->check(dabgpu_msc_add_subchannel(m_ctx->get(), 0, &subs[i], &index))) { m_rejected_ids.push_back(ids[i]); continue; }
            auto ch = std::make_unique<Channel>();
            ch->subchannel.id = ids[i];
            ch->subchannel.start_address = uint16_t(subs[i].start_address);
            ch->subchannel.length = uint16_t(subs[i].length);
            ch->subchannel.is_uep = subs[i].is_uep != 0;
            ch->subchannel.uep_prot_index = uint8_t(subs[i].uep_prot_index);
            ch->subchannel.eep_prot_level = uint8_t(subs[i].eep_prot_level);
            ch->subchannel.eep_type = subs[i].eep_type_b ? EEP_Type::TYPE_B : EEP_Type::TYPE_A;
            ch->is_dabplus = subs[i].is_dabplus != 0;
            if (size_t(index) != m_channels.size()) continue;   // cannot happen while this object is the only writer of the table
            m_channels.push_back(std::move(ch));
            m_configured_ids.push_back(ids[i]);
            m_obs_audio_channel.Notify(ids[i], *m_channels.back());
        }
    }
    void deliver() {
        dabgpu_chan_status st;
        if (!m_ctx->check(dabgpu_chan_get_status(m_ctx->get(), 0, &st)) || !st.decoded) return;
        if (!m_ctx->check(dabgpu_chan_get_fic(m_ctx->get(), 0, m_fibs.data(), m_crc.data()))) return;
        for (int i = 0; i < m_params.nb_fibs; i++)
            if (m_crc[size_t(i)]) {
                m_obs_on_fib.Notify(span<const uint8_t>(m_fibs.data() + 32 * size_t(i), 30));
                if (m_self_configure) m_autocfg.ProcessFIB(m_fibs.data() + 32 * size_t(i), 30);
            }
        for (size_t k = 0; k < m_channels.size(); k++) {
            Channel& ch = *m_channels[k];
            uint8_t valid[8] = {0};
            int nb = 0;
            if (!m_ctx->check(dabgpu_chan_get_msc(m_ctx->get(), 0, int(k), m_msc.data(), m_msc.size(), valid, &nb))) continue;
            for (int c = 0; c < m_params.nb_cifs; c++)
                if (valid[c]) ch.on_msc_data.Notify(span<const uint8_t>(m_msc.data() + size_t(c) * size_t(nb), size_t(nb)));
            if (ch.is_dabplus) {
                size_t n = 0;
                if (m_ctx->check(dabgpu_chan_get_dabplus_events(m_ctx->get(), 0, int(k), m_log.data(), m_log.size(), &n))) ch.dabplus.dispatch(m_log.data(), n);
            }
        }
        if (m_self_configure) update_after_processing();
    }
public:
    explicit BasicRadio(const DAB_Parameters& params, const size_t nb_threads = 0, std::shared_ptr<Context> ctx = nullptr)
        : m_params(params), m_ctx(ctx), m_fibs(size_t(params.nb_fibs) * 32), m_crc(size_t(params.nb_fibs)),
          m_msc(size_t(params.nb_cifs) * DABGPU_CIF_OUT_STRIDE), m_log(1 << 16) {
        (void)nb_threads;
        if (!m_ctx) {
            const int mode = params.nb_cifs == 4 ? 1 : params.nb_cifs == 2 ? 4 : (params.nb_fibs_per_cif == 4 ? 3 : 2);
            m_ctx = std::make_shared<Context>(mode, DABGPU_IQ_C32);
        }
    }
    // (re)declares the sub-channels to decode; de-interleaver and superframe state start empty
    bool SetSubchannels(const std::vector<std::pair<Subchannel, bool>>& subs) {
        std::lock_guard<std::mutex> lock(m_ctx->mutex());
        std::vector<dabgpu_subchannel> d;
        m_channels.clear();
        for (const auto& s : subs) {
            d.push_back(to_abi(s.first, s.second));
            auto ch = std::make_unique<Channel>();
            ch->subchannel = s.first;
            ch->is_dabplus = s.second;
            m_channels.push_back(std::move(ch));
        }
        return m_ctx->check(dabgpu_msc_configure(m_ctx->get(), 0, d.data(), int(d.size())));
    }
    // must be exactly nb_frame_bits, otherwise the frame is dropped like in the reference (basic_radio.cpp:41-46)
    void Process(span<const viterbi_bit_t> buf) {
        if (int(buf.size()) != m_params.nb_frame_bits) return;
        std::lock_guard<std::mutex> lock(m_ctx->mutex());
        if (!m_ctx->check(dabgpu_softbits_push(m_ctx->get(), buf.data(), buf.size(), 0, 1))) return;
        if (!m_ctx->check(dabgpu_chan_decode(m_ctx->get(), 0, 1))) return;
        deliver();
    }
    // Radio_Block: decode every frame the OFDM stage of the same context left in the device ring (no host round trip).
    // The caller holds the context mutex.
    void ProcessDeviceFrames() {
        for (;;) {
            if (!m_ctx->check(dabgpu_chan_decode(m_ctx->get(), 0, 1))) return;
            dabgpu_chan_status st;
            if (!m_ctx->check(dabgpu_chan_get_status(m_ctx->get(), 0, &st)) || !st.decoded) return;
            deliver();
        }
    }
    // Let the FIC decide: FIG 0/1 + 0/2 (+ 0/3, 0/14) are parsed on the host (fic_autoconfig.hpp) and the sub-channel set is
    // applied after the frame that completed it, like the reference's BasicRadio does without being asked.
    void EnableSelfConfiguration(bool on = true) { m_self_configure = on; }
    auto& On_Audio_Channel() { return m_obs_audio_channel; }
    const FIC_Autoconfig& GetFICDatabase() const { return m_autocfg; }
    Channel* Get_Channel(size_t index) { return index < m_channels.size() ? m_channels[index].get() : nullptr; }
    size_t GetTotalChannels() const { return m_channels.size(); }
    auto& GetMutex() { return m_mutex_data; }
    auto& On_FIB() { return m_obs_on_fib; }
    size_t GetTotalThreads() const { return 1; }
    const std::string& LastError() const { return m_ctx->last_error(); }
};

// ---------------------------------------------------------------------------------------------------------------
// Radio_Block (src/radio_block.h): owns the demodulator and the radio of one tuner.  Both halves share one GPU
// context, so the 230400-byte soft-bit frame the reference pushes through a ThreadedRingBuffer between two threads
// (src/radio_block.cpp:20-44) never leaves the device.
// ---------------------------------------------------------------------------------------------------------------
constexpr int TRANSMISSION_MODE = 1;   // src/radio_block.cpp:9

class Radio_Block {
    const OFDM_Params m_ofdm_params;
    const DAB_Parameters m_dab_params;
    const size_t m_ofdm_total_threads, m_dab_total_threads;
    std::shared_ptr<Context> m_ctx;
    std::shared_ptr<OFDM_Demod> m_ofdm_demodulator;
    std::mutex m_mutex_basic_radio;
    std::shared_ptr<BasicRadio> m_basic_radio;
public:
    Radio_Block(size_t ofdm_total_threads, size_t dab_total_threads, int transmission_mode = TRANSMISSION_MODE)
        : m_ofdm_params(get_DAB_OFDM_params(transmission_mode)), m_dab_params(get_dab_parameters(transmission_mode)),
          m_ofdm_total_threads(ofdm_total_threads), m_dab_total_threads(dab_total_threads),
          m_ctx(std::make_shared<Context>(transmission_mode, DABGPU_IQ_C32)) {
        m_ofdm_demodulator = std::make_shared<OFDM_Demod>(m_ofdm_params, span<const std::complex<float>>(), span<const int>(),
                                                          int(m_ofdm_total_threads), m_ctx);
        m_ofdm_demodulator->SetDeviceFrameHook([this]() {
            std::shared_ptr<BasicRadio> radio;
            {
                std::lock_guard<std::mutex> lock(m_mutex_basic_radio);
                radio = m_basic_radio;
            }
            if (radio) radio->ProcessDeviceFrames();
        });
        reset_radio();
    }
    void reset_radio() {
        auto radio = std::make_shared<BasicRadio>(m_dab_params, m_dab_total_threads, m_ctx);
        std::lock_guard<std::mutex> lock(m_mutex_basic_radio);
        m_basic_radio = radio;
    }
    std::shared_ptr<OFDM_Demod> get_ofdm_demodulator() { return m_ofdm_demodulator; }
    std::shared_ptr<BasicRadio> get_basic_radio() {
        std::lock_guard<std::mutex> lock(m_mutex_basic_radio);
        return m_basic_radio;
    }
};

}  // namespace dabgpu_host
