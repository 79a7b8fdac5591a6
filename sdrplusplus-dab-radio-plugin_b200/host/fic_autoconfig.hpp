// fic_autoconfig.hpp -- host side of the C ABI: turns the FIBs that dabgpu_chan_get_fic / FIC_Decoder::OnFIB deliver into the
// sub-channel table dabgpu_msc_configure wants, so that a receiver configures itself from the Fast Information Channel
// instead of being told its sub-channels.  Header only, C++17, no CUDA.
//
// Restates, for the FIGs that define the multiplex configuration of the MSC, the reference's
//   FIG_Processor::ProcessFIB / Type_0 / Ext_1 / Ext_2 / Ext_3 / Ext_14   dab/fic/fig_processor.cpp:94-152, 154-193, 302-362, 364-487, 489-551, ext 14
//   Radio_FIG_Handler::OnSubchannel_1_{Short,Long}, OnServiceComponent_1_*, OnServiceComponent_2_PacketDataType, OnSubchannel_2_FEC
//                                                                          dab/radio_fig_handler.cpp:35-75, 77-177, 179-214, 474-481
//   DatabaseEntityUpdater::UpdateField (first value wins, a different later value is a conflict and is dropped)
//                                                                          dab/database/dab_database_updater.h:89-104
//   SubchannelUpdater / ServiceComponentUpdater required-field masks      dab/database/dab_database_updater.cpp:94-107, 161-169, 172-180, 222-228
//   BasicRadio::UpdateAfterProcessing (which sub-channels get a decoder)   basic_radio/basic_radio.cpp:83-154
// Every other FIG (labels, date/time, linkage, 0/8 ...) is skipped by its length byte.  A secondary component that FIG 0/2 lists
// before any FIG 0/8 has created it is ignored, exactly as the reference ignores it at that point.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <vector>
#include "../../include/dabgpu.h"

namespace dabgpu_host {

struct FicSubchannel {
    uint8_t id = 0;
    uint16_t start_address = 0, length = 0;
    bool is_uep = false;
    uint8_t uep_prot_index = 0, eep_prot_level = 0;
    uint8_t eep_type = 0xFF;     // 0 = A, 1 = B, 0xFF undefined
    uint8_t fec_scheme = 0xFF;   // FIG 0/14, 0xFF undefined
    uint8_t dirty = 0;
    bool is_complete = false;
};

struct FicServiceComponent {
    uint32_t service_value = 0;
    uint8_t service_type = 0xFF;   // 0 = 16 bit SId, 1 = 32 bit SId
    uint8_t component_id = 0;
    uint16_t global_id = 0xFFFF;
    uint8_t subchannel_id = 0;
    uint16_t packet_address = 0;
    uint8_t transport_mode = 0xFF;   // 0 stream audio, 1 stream data, 3 packet data
    uint8_t audio_type = 0xFF;       // 0 DAB (MPEG-1 layer II), 63 DAB+
    uint8_t data_type = 0xFF;
    uint16_t dirty = 0;
    bool is_complete = false;
    uint32_t service_uuid() const { return service_type == 0 ? (service_value & 0xFFFFu) : service_value; }
};

class FIC_Autoconfig {
public:
    // One CRC-checked FIB (30 bytes, the CRC itself excluded).  Returns true when the database changed.
    bool ProcessFIB(const uint8_t* buf, int N = 30) {
        const uint64_t before = m_updates;
        int cur = 0;
        while (cur < N) {
            const uint8_t header = buf[cur];
            if (header == 0xFF) break;                       // end marker
            const int type = header >> 5, len = header & 31;
            if (len + 1 > N - cur) break;                    // "fig specified length overflows buffer"
            const uint8_t* fig = buf + cur + 1;
            cur += len + 1;
            if (type == 0) ProcessType0(fig, len);
            else if (type == 1 || type == 2 || type == 6) continue;   // labels, conditional access: not needed here
            else break;                                      // type 7 ends the FIB, types 3..5 are invalid: the reference stops too
        }
        return m_updates != before;
    }

    const std::vector<FicSubchannel>& subchannels() const { return m_subchannels; }
    const std::vector<FicServiceComponent>& components() const { return m_components; }
    uint64_t total_updates() const { return m_updates; }

    // The sub-channels BasicRadio would attach a decoder to, in database order: complete sub-channel, first service component
    // that names it complete, stream audio DAB+ or DAB (basic_radio.cpp:97-141).  ids[i] is the SubChId of out[i].
    void Runnable(std::vector<dabgpu_subchannel>& out, std::vector<uint8_t>& ids) const {
        out.clear();
        ids.clear();
        for (const FicSubchannel& sc : m_subchannels) {
            if (!sc.is_complete) continue;
            const FicServiceComponent* comp = nullptr;
            for (const FicServiceComponent& c : m_components)
                if (c.subchannel_id == sc.id) { comp = &c; break; }
            if (!comp || !comp->is_complete) continue;
            if (comp->transport_mode != 0 || (comp->audio_type != 63 && comp->audio_type != 0)) continue;
            dabgpu_subchannel d;
            d.start_address = sc.start_address;
            d.length = sc.length;
            d.is_uep = sc.is_uep ? 1 : 0;
            d.uep_prot_index = sc.uep_prot_index;
            d.eep_prot_level = sc.eep_prot_level;
            d.eep_type_b = (sc.eep_type == 1) ? 1 : 0;
            d.is_dabplus = (comp->audio_type == 63) ? 1 : 0;
            out.push_back(d);
            ids.push_back(sc.id);
        }
    }

private:
    enum : uint8_t { SC_START = 0x80, SC_LENGTH = 0x40, SC_IS_UEP = 0x20, SC_UEP_INDEX = 0x10, SC_EEP_LEVEL = 0x08, SC_EEP_TYPE = 0x04, SC_FEC = 0x02,
                     SC_REQ_UEP = 0xF0, SC_REQ_EEP = 0xEC };
    enum : uint16_t { CO_TRANSPORT = 0x100, CO_AUDIO = 0x080, CO_DATA = 0x040, CO_SUBCHANNEL = 0x020, CO_GLOBAL_ID = 0x010, CO_PACKET_ADDR = 0x004,
                      CO_REQ_AUDIO = 0x1A0, CO_REQ_DATA = 0x160, CO_REQ_PACKET = 0x165 };
    // Sub-channel size in capacity units per UEP table index, as the reference lists them (subchannel_protection_tables.h:21-86,
    // column 0).  Rows 33 and 34 (128 kbit/s, levels 5 and 4) are 84 and 64 there -- EN 300 401 table 8 has 64 and 84 --
    // and a drop-in must size the sub-channel like the code it replaces.
    static int uep_size(int index) {
        static const uint16_t S[64] = {16, 21, 24, 29, 35, 24, 29, 35, 42, 52, 29, 35, 42, 52, 32, 42, 48, 58, 70, 40, 52, 58,
                                       70, 84, 48, 58, 70, 84, 104, 58, 70, 84, 104, 84, 64, 96, 116, 140, 80, 104, 116, 140, 168, 96,
                                       116, 140, 168, 208, 116, 140, 168, 208, 232, 128, 168, 192, 232, 280, 160, 208, 280, 192, 280, 416};
        return S[index];
    }

    template <typename F, typename V, typename M> bool set(F& dst, V src, M& dirty, M flag) {   // UpdateField
        if (dirty & flag) return dst == F(src) ? true : false;   // unchanged, or a conflict: the first value stays
        dirty = M(dirty | flag);
        dst = F(src);
        m_updates++;
        return true;
    }
    FicSubchannel& subchannel(uint8_t id) {
        for (FicSubchannel& s : m_subchannels) if (s.id == id) return s;
        m_subchannels.emplace_back();
        m_subchannels.back().id = id;
        m_updates++;
        return m_subchannels.back();
    }
    static void finish(FicSubchannel& s) {
        s.is_complete = s.is_uep ? ((s.dirty & SC_REQ_UEP) == SC_REQ_UEP) : ((s.dirty & SC_REQ_EEP) == SC_REQ_EEP);
    }
    static void finish(FicServiceComponent& c) {
        const uint16_t req = c.transport_mode == 0 ? CO_REQ_AUDIO : (c.transport_mode == 1 ? CO_REQ_DATA : CO_REQ_PACKET);
        c.is_complete = (c.dirty & req) == req;
    }
    FicServiceComponent* component_primary(uint32_t value, uint8_t type) {   // GetServiceComponentUpdater_Service(service_id, 0)
        const uint32_t uuid = type == 0 ? (value & 0xFFFFu) : value;
        for (FicServiceComponent& c : m_components) if (c.service_uuid() == uuid && c.component_id == 0) return &c;
        m_components.emplace_back();
        FicServiceComponent& c = m_components.back();
        c.service_value = value; c.service_type = type; c.component_id = 0;
        m_updates++;
        return &c;
    }
    FicServiceComponent* component_by_subchannel(uint32_t value, uint8_t type, uint8_t subchannel_id) {
        const uint32_t uuid = type == 0 ? (value & 0xFFFFu) : value;
        for (FicServiceComponent& c : m_components) if (c.service_uuid() == uuid && c.subchannel_id == subchannel_id) return &c;
        return nullptr;
    }
    FicServiceComponent* component_by_global_id(uint16_t gid) {
        for (FicServiceComponent& c : m_components) if (c.global_id == gid) return &c;
        return nullptr;
    }
    void set_audio_type(FicServiceComponent& c, uint8_t v) { if (!(c.dirty & CO_DATA)) set(c.audio_type, v, c.dirty, uint16_t(CO_AUDIO)); }
    void set_data_type(FicServiceComponent& c, uint8_t dscty) {
        if (dscty != 5 && dscty != 24 && dscty != 60 && dscty != 63) return;   // "Unsupported data service type"
        if (!(c.dirty & CO_AUDIO)) set(c.data_type, dscty, c.dirty, uint16_t(CO_DATA));
    }

    void ProcessType0(const uint8_t* buf, int n) {
        if (n < 1) return;
        const bool pd = (buf[0] >> 5) & 1;
        const int ext = buf[0] & 31;
        const uint8_t* b = buf + 1;
        n -= 1;
        if (ext == 1) Ext1(b, n);
        else if (ext == 2) Ext2(b, n, pd);
        else if (ext == 3) Ext3(b, n);
        else if (ext == 14) Ext14(b, n);
    }

    void Ext1(const uint8_t* buf, int N) {   // sub-channel organisation
        int cur = 0;
        while (cur < N) {
            const uint8_t* d = buf + cur;
            const int remain = N - cur;
            if (remain < 3) break;
            const uint8_t id = d[0] >> 2;
            const uint16_t start = uint16_t((d[0] & 3) << 8) | d[1];
            const bool is_long = d[2] >> 7;
            const int nb = is_long ? 4 : 3;
            if (nb > remain) break;
            FicSubchannel& s = subchannel(id);
            if (!is_long) {
                const bool table_switch = (d[2] >> 6) & 1;
                const int index = d[2] & 63;
                set(s.start_address, start, s.dirty, uint8_t(SC_START));
                set(s.is_uep, true, s.dirty, uint8_t(SC_IS_UEP));
                if (!table_switch) {
                    if (set(s.is_uep, true, s.dirty, uint8_t(SC_IS_UEP))) set(s.uep_prot_index, index, s.dirty, uint8_t(SC_UEP_INDEX));
                    set(s.length, uep_size(index), s.dirty, uint8_t(SC_LENGTH));
                }
            } else {
                const int option = (d[2] >> 4) & 7, level = (d[2] >> 2) & 3;
                const uint16_t size = uint16_t((d[2] & 3) << 8) | d[3];
                set(s.is_uep, false, s.dirty, uint8_t(SC_IS_UEP));
                set(s.start_address, start, s.dirty, uint8_t(SC_START));
                if (set(s.is_uep, false, s.dirty, uint8_t(SC_IS_UEP))) set(s.eep_type, option ? 1 : 0, s.dirty, uint8_t(SC_EEP_TYPE));
                if (set(s.is_uep, false, s.dirty, uint8_t(SC_IS_UEP))) set(s.eep_prot_level, level, s.dirty, uint8_t(SC_EEP_LEVEL));
                set(s.length, size, s.dirty, uint8_t(SC_LENGTH));
            }
            finish(s);
            cur += nb;
        }
    }

    void Ext2(const uint8_t* buf, int N, bool pd) {   // service organisation
        const int nb_sid = pd ? 4 : 2, nb_header = nb_sid + 1;
        int cur = 0;
        while (cur < N) {
            const uint8_t* sb = buf + cur;
            const int remain = N - cur;
            if (nb_header > remain) return;
            uint32_t sid;
            if (pd) sid = uint32_t(sb[0]) << 24 | uint32_t(sb[1]) << 16 | uint32_t(sb[2]) << 8 | sb[3];
            else sid = uint32_t(sb[0]) << 8 | sb[1];
            const uint8_t stype = pd ? 1 : 0;
            const int n_comp = sb[nb_sid] & 15;
            const int total = 2 * n_comp + nb_header;
            if (total > remain) return;
            for (int i = 0; i < n_comp; i++) {
                const uint8_t b0 = sb[nb_header + 2 * i], b1 = sb[nb_header + 2 * i + 1];
                const int tmid = b0 >> 6;
                const bool primary = (b1 >> 1) & 1;
                if (tmid == 0 || tmid == 1) {
                    const uint8_t ty = b0 & 63, sub = b1 >> 2;
                    FicServiceComponent* c = primary ? component_primary(sid, stype) : component_by_subchannel(sid, stype, sub);
                    if (!c) continue;
                    set(c->subchannel_id, sub, c->dirty, uint16_t(CO_SUBCHANNEL));
                    set(c->transport_mode, tmid, c->dirty, uint16_t(CO_TRANSPORT));
                    if (tmid == 0) { if (ty == 0 || ty == 63) set_audio_type(*c, ty); }
                    else set_data_type(*c, ty);
                    finish(*c);
                } else if (tmid == 3) {
                    const uint16_t scid = uint16_t((b0 & 63) << 6) | (b1 >> 2);
                    FicServiceComponent* c = primary ? component_primary(sid, stype) : component_by_global_id(scid);
                    if (!c) continue;
                    set(c->transport_mode, 3, c->dirty, uint16_t(CO_TRANSPORT));
                    set(c->global_id, scid, c->dirty, uint16_t(CO_GLOBAL_ID));
                    finish(*c);
                } else {
                    return;   // reserved TMId: the reference abandons the FIG
                }
            }
            cur += total;
        }
    }

    void Ext3(const uint8_t* buf, int N) {   // service component in packet mode
        int cur = 0;
        while (cur < N) {
            const int remain = N - cur;
            if (remain < 5) return;
            const uint8_t* b = buf + cur;
            const uint16_t scid = uint16_t(b[0]) << 4 | (b[1] >> 4);
            const bool caorg = b[1] & 1;
            const uint8_t dscty = b[2] & 63, sub = b[3] >> 2;
            const uint16_t addr = uint16_t((b[3] & 3) << 8) | b[4];
            const int len = caorg ? 7 : 5;
            if (len > remain) return;
            if (FicServiceComponent* c = component_by_global_id(scid)) {
                set(c->subchannel_id, sub, c->dirty, uint16_t(CO_SUBCHANNEL));
                set(c->transport_mode, 3, c->dirty, uint16_t(CO_TRANSPORT));
                set(c->global_id, scid, c->dirty, uint16_t(CO_GLOBAL_ID));
                set(c->packet_address, addr, c->dirty, uint16_t(CO_PACKET_ADDR));
                set_data_type(*c, dscty);
                finish(*c);
            }
            cur += len;
        }
    }

    void Ext14(const uint8_t* buf, int N) {   // FEC sub-channel organisation
        for (int i = 0; i < N; i++) {
            FicSubchannel& s = subchannel(buf[i] >> 2);
            set(s.fec_scheme, buf[i] & 3, s.dirty, uint8_t(SC_FEC));
            finish(s);
        }
    }

    std::vector<FicSubchannel> m_subchannels;
    std::vector<FicServiceComponent> m_components;
    uint64_t m_updates = 0;
};

}  // namespace dabgpu_host
