// vgl_host.hpp -- C++ host side above the C ABI (include/vgl.h), mirroring the reference's operator
// interface for the hot path so that the reference's driver loop can switch over with a few lines:
//
//   reference                                   here
//   ---------                                   ----
//   simRecord (bcf_utils.h:81-396)              vgl::SimRecordView  (same member names, read-only views)
//   simulate_record_values(sim) per site        vgl::BatchSimulator::push_site() + on_record callback
//     (vcfgl.cpp:327, called :1522,1552,1611)
//   return codes 0 / -3 / -4                    SimRecordView::ret
//   sim->add_tags()  (bcf_utils.cpp:426-507)    the callback passes the view's arrays to bcf_update_*
//   ERROR()/exit(1)  (shared.h:292-327)         vgl::Error exception carrying the vgl_status
//
// Header-only; link with -lvgl.  One BatchSimulator per GPU (one host thread each).
#pragma once
#include "../../include/vgl.h"

#include <cstdint>
#include <cstring>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace vgl {

struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string& what) : std::runtime_error(what), status(s) {}
};

// One site's results with the reference's simRecord field names (bcf_utils.h:85-211).
struct SimRecordView {
    int64_t site_id = 0;  // global running index given to push_site()
    void* user = nullptr; // whatever the caller attached to the site (e.g. its bcf1_t*)
    int ret = 0;          // 0, -3 (simulated invariant, vcfgl.cpp:677), -4 (empty, vcfgl.cpp:401)
    int nSamples = 0, nAlleles = 0, nAllelesObserved = 0, nGenotypes = 0;
    int allele_unobserved = -1;
    int alleles2acgt[5], acgt2alleles[5];
    std::string alleles;  // "A,C,<*>" ... as passed to bcf_update_alleles_str (vcfgl.cpp:739-782)
    const int32_t* fmt_dp_arr = nullptr; // [nSamples]
    int32_t info_dp_arr[1] = {0};
    const float* gl_arr = nullptr;       // [nSamples*nGenotypes]
    const int32_t* pl_arr = nullptr;
    const float* gp_arr = nullptr;
    const int32_t *fmt_ad_arr = nullptr, *fmt_adf_arr = nullptr, *fmt_adr_arr = nullptr; // [nSamples*nAlleles]
    const int32_t *info_ad_arr = nullptr, *info_adf_arr = nullptr, *info_adr_arr = nullptr; // [nAlleles]
    const float* qs_arr = nullptr;       // [nAlleles]
    const float* i16_arr = nullptr;      // [16]
    // current_size_bcf_tag_number[] equivalents (bcf_utils.h:25-36)
    int size_fmt_G() const { return nSamples * nGenotypes; }
    int size_fmt_R() const { return nSamples * nAlleles; }
};

// ---------------------------------------------------------------------------------------------------------------
// gVCF (-doGVCF 1): the record sequence of a run from the device's per-batch merge (vgl_gvcf_merge).  What is left of
// prepare_gvcf_block() (bcf_utils.cpp:662-942) on the host is the seam between batches: a batch's first block continues
// the block still open from the batch before under the reference's own three conditions (same contig, pos <= end + 1,
// same dp range: bcf_utils.cpp:711, 719, 790), with the same minima (:838-866).
//
//   reference                                            here
//   ---------                                            ----
//   write_record_values(sim) per site (vcfgl.cpp:165)    feed() once per batch, after vgl_wait + vgl_gvcf_merge
//   GVCF_WRITE_SIMREC                                    on_site(global site index): write the site's own record
//   GVCF_FLUSH_BLOCK + grec (bcf_utils.cpp:876-905)      on_block(Block): alleles / QS from site `first`, END, MIN_DP, DP, PL
//   write_record_values(NULL) at the end (vcfgl.cpp:169) finish()
class GvcfStitcher {
public:
    struct Block {
        int64_t first = 0;   // global index of the founder site
        int32_t rid = 0;
        int64_t start = 0, end = 0; // 0-based positions of the first / last member
        int32_t min_dp = 0, dp_range = 0, n_members = 0;
        std::vector<int32_t> dp, pl; // [S], [S * 3] (empty without PL)
        std::string alleles;         // founder's allele string and INFO/QS, filled by the annotate hook of feed()
        std::vector<float> qs;
        void* user = nullptr;        // founder's user pointer
    };
    GvcfStitcher(int n_samples, std::function<void(int64_t)> on_site, std::function<void(const Block&)> on_block)
        : S_(n_samples), on_site_(std::move(on_site)), on_block_(std::move(on_block))
    {
    }

    // annotate(block, founder's site index in this batch): called when a block is opened, to copy what the block record takes
    // from its founder (bcf_utils.cpp:817-832) while the batch is still in memory
    void feed(const vgl_gvcf_out& o, const vgl_gvcf_site_in* sites, int32_t n_sites, const std::function<void(Block&, int32_t)>& annotate = nullptr)
    {
        for (int32_t k = 0; k < o.n_recs; ++k) {
            const vgl_gvcf_rec& r = o.recs[k];
            if (r.n_members == 0) {
                flush();
                on_site_(base_ + r.first_site);
                continue;
            }
            const vgl_gvcf_site_in &f = sites[r.first_site], &l = sites[r.last_site];
            const int32_t* dp = o.dp + (size_t)r.plane * S_;
            const int32_t* pl = o.pl ? o.pl + (size_t)r.plane * S_ * 3 : nullptr;
            if (open_ && k == 0 && cur_.rid == f.rid && f.pos <= cur_.end + 1 && cur_.dp_range == r.dp_range) {
                cur_.end = l.pos;
                cur_.min_dp = std::min(cur_.min_dp, r.min_dp);
                cur_.n_members += r.n_members;
                for (int s = 0; s < S_; ++s) {
                    cur_.dp[s] = std::min(cur_.dp[s], dp[s]);
                    if (pl && !cur_.pl.empty()) {
                        int32_t* g = &cur_.pl[3 * (size_t)s];
                        const int32_t* m = pl + 3 * (size_t)s;
                        if (m[1] < g[1] || (m[1] == g[1] && m[2] < g[2])) g[1] = m[1], g[2] = m[2];
                    }
                }
                continue;
            }
            flush();
            cur_.first = base_ + r.first_site;
            cur_.rid = f.rid, cur_.start = f.pos, cur_.end = l.pos;
            cur_.min_dp = r.min_dp, cur_.dp_range = r.dp_range, cur_.n_members = r.n_members;
            cur_.dp.assign(dp, dp + S_);
            if (pl) cur_.pl.assign(pl, pl + 3 * (size_t)S_);
            else cur_.pl.clear();
            if (annotate) annotate(cur_, r.first_site);
            open_ = true;
        }
        base_ += n_sites;
    }
    void finish() { flush(); }

private:
    void flush()
    {
        if (open_) on_block_(cur_);
        open_ = false;
    }
    int S_;
    std::function<void(int64_t)> on_site_;
    std::function<void(const Block&)> on_block_;
    Block cur_;
    bool open_ = false;
    int64_t base_ = 0;
};

// allele string of a simulated site; no-reads sites follow simulate_site_with_no_reads (vcfgl.cpp:228-315)
inline std::string alleles_string(const vgl_site_out& s, int do_unobserved, int do_gvcf)
{
    const char* nonref = (do_unobserved == 2 || do_unobserved == 5) ? "<NON_REF>" : "<*>";
    if (s.info_dp == 0) {
        if (do_gvcf) return "<NON_REF>";
        switch (do_unobserved) {
        case 0: return ".";
        case 1: return "<*>";
        case 2: return "<NON_REF>";
        case 3: return "A,C,G,T";
        case 4: return "A,C,G,T,<*>";
        default: return "A,C,G,T,<NON_REF>";
        }
    }
    std::string out;
    for (int a = 0; a < s.n_alleles; ++a) {
        if (a) out += ',';
        const int b = s.alleles2acgt[a];
        if (b == 4) out += nonref;
        else out += "ACGT"[b];
    }
    return out;
}

class BatchSimulator {
public:
    using Callback = std::function<void(const SimRecordView&)>;

    // params: fill from the parsed argStruct (io.h:40-148); host_output is forced on.  With VGL_HOST_NARROW the
    // integer planes cross PCIe as 8/16-bit values and are widened here, one site at a time, into the int32
    // arrays bcf_update_format_int32 takes (it narrows them again itself, htslib/vcf.c:2249-2294).
    BatchSimulator(vgl_params params, Callback on_record) : BatchSimulator(params, std::vector<int>{params.device_id}, std::move(on_record)) {}

    // Several GPUs of one box: one context per device, consecutive batches -- contiguous ranges of the global
    // site index -- go to the devices in turn, and the results come back through the one callback strictly in site order (the
    // ordered host-side merge; the gVCF block machine runs over the merged stream, so blocks cross device boundaries like
    // batch boundaries).  Every draw is keyed by the global site index, so the records are the same bytes for any device list.
    // A device may be listed more than once (several contexts on one GPU).
    BatchSimulator(vgl_params params, const std::vector<int>& device_ids, Callback on_record) : prm_(params), cb_(std::move(on_record))
    {
        if (device_ids.empty()) throw Error(VGL_EINVAL, "BatchSimulator: empty device list");
        prm_.abi_version = VGL_ABI_VERSION;
        if (prm_.host_output != VGL_HOST_NARROW) prm_.host_output = VGL_HOST_I32;
        if (prm_.n_slots < 2) prm_.n_slots = 2;
        for (int d : device_ids) {
            vgl_params q = prm_;
            q.device_id = d;
            vgl_ctx* c = nullptr;
            const int rc = vgl_create(&q, &c);
            if (rc != VGL_OK) {
                for (vgl_ctx* x : ctxs_) vgl_destroy(x);
                throw Error(rc, std::string("vgl_create: ") + vgl_strerror(rc));
            }
            ctxs_.push_back(c);
        }
        ctx_ = ctxs_[0];
        n_lanes_ = (int)ctxs_.size() * prm_.n_slots;
        pending_.resize(n_lanes_);
        open_slot();
    }
    ~BatchSimulator() { for (vgl_ctx* c : ctxs_) vgl_destroy(c); }
    int n_devices() const { return (int)ctxs_.size(); }
    BatchSimulator(const BatchSimulator&) = delete;
    BatchSimulator& operator=(const BatchSimulator&) = delete;

    // In place of simulate_record_values(sim): true_gts_acgt_int is what check_rec_alleles() filled
    // (vcfgl.cpp:133-146): 2*nSamples ints in {-1,0,1,2,3}.  Results arrive later through the callback,
    // strictly in push order.
    void push_site(const int* true_gts_acgt_int, void* user = nullptr)
    {
        uint8_t* row = in_ + (size_t)fill_ * prm_.n_samples;
        for (int s = 0; s < prm_.n_samples; ++s) {
            const int h0 = true_gts_acgt_int[2 * s], h1 = true_gts_acgt_int[2 * s + 1];
            row[s] = VGL_GT_PACK(h0 < 0 ? VGL_GT_MISSING : h0, h1 < 0 ? VGL_GT_MISSING : h1);
        }
        pending_[cur_].users.push_back(user);
        if (++fill_ == prm_.max_batch_sites) submit_current();
    }

    // Input path (include/vgl.h): a whole batch whose genotypes were parsed on the device by vgl_parse_vcf().  Site k takes the
    // row of parsed record row_map[k], or fill_gt in every sample when row_map[k] < 0 (an -explode 1 site).  Not to be mixed
    // with push_site() inside one batch: a partially filled host batch is submitted first.
    void push_device_sites(vgl_parser* ps, const int32_t* row_map, int32_t n, uint8_t fill_gt, void* const* users = nullptr)
    {
        if (ctxs_.size() != 1) throw Error(VGL_EINVAL, "push_device_sites: the parser belongs to one device; use one context");
        if (fill_ > 0) submit_current();
        check(vgl_place_rows(ctx_, cur_, ps, row_map, 0, n, fill_gt), "vgl_place_rows");
        Pending& p = pending_[cur_];
        p.users.assign((size_t)n, nullptr);
        if (users) p.users.assign(users, users + n);
        p.first = next_site_;
        check(vgl_submit(ctx_, cur_, next_site_, n, nullptr, VGL_SUBMIT_GT_ON_DEVICE), "vgl_submit");
        p.in_flight = true;
        next_site_ += n;
        cur_ = (cur_ + 1) % n_lanes_;
        open_slot();
    }
    vgl_ctx* context() { return ctx_; }
    const vgl_params& params() const { return prm_; }

    // -doGVCF 1 (write_record_values + prepare_gvcf_block, vcfgl.cpp:165-207, bcf_utils.cpp:662-942): every finished batch is
    // merged on the device (vgl_gvcf_merge) and stitched to the batch before it.  `where(user)` gives a site's (rid, pos).
    // Sites that end up inside a block are not delivered through the record callback; the block is, through on_block, in
    // output order between the regular records (its alleles / qs / user are the founder's).  Call before the first push.
    void enable_gvcf(std::vector<int32_t> gvcf_dps, std::function<vgl_gvcf_site_in(void*)> where,
                     std::function<void(const GvcfStitcher::Block&)> on_block)
    {
        gvcf_dps_ = std::move(gvcf_dps);
        where_ = std::move(where);
        stitch_.reset(new GvcfStitcher(prm_.n_samples, [this](int64_t g) { deliver(*cur_out_, *cur_pending_, (int)(g - cur_pending_->first)); },
                                       std::move(on_block)));
    }

    // end of input: run the partial batch and deliver everything outstanding
    void finish()
    {
        if (fill_ > 0) submit_current();
        for (int k = 0; k < n_lanes_; ++k) drain((cur_ + k) % n_lanes_);
        if (stitch_) stitch_->finish(); // the block still open at the end (vcfgl.cpp:169-177)
    }

    int64_t sites_pushed() const { return next_site_; }

private:
    struct Pending {
        bool in_flight = false;
        int64_t first = 0;
        std::vector<void*> users;
    };

    void check(int rc, const char* what)
    {
        if (rc != VGL_OK) throw Error(rc, std::string(what) + ": " + vgl_strerror(rc) + " (" + vgl_last_error(ctx_) + ")");
    }
    // lane = (device, slot): consecutive batches take consecutive lanes, i.e. alternate between the devices
    vgl_ctx* ctx_of(int lane) const { return ctxs_[(size_t)lane % ctxs_.size()]; }
    int slot_of(int lane) const { return lane / (int)ctxs_.size(); }

    void open_slot()
    {
        drain(cur_); // the slot we are about to refill must have been delivered
        int64_t cap = 0;
        check(vgl_input_buffer(ctx_of(cur_), slot_of(cur_), &in_, &cap), "vgl_input_buffer");
        fill_ = 0;
        pending_[cur_].users.clear();
    }

    void submit_current()
    {
        Pending& p = pending_[cur_];
        p.first = next_site_;
        check(vgl_submit(ctx_of(cur_), slot_of(cur_), next_site_, fill_, nullptr, 0), "vgl_submit");
        p.in_flight = true;
        next_site_ += fill_;
        cur_ = (cur_ + 1) % n_lanes_;
        open_slot(); // delivers the oldest batch while this one runs on its GPU
    }

    void drain(int slot)
    {
        Pending& p = pending_[slot];
        if (!p.in_flight) return;
        vgl_batch_out out;
        check(vgl_wait(ctx_of(slot), slot_of(slot), &out), "vgl_wait");
        if (out.status != VGL_OK) throw Error(out.status, vgl_strerror(out.status));
        if (stitch_) {
            std::vector<vgl_gvcf_site_in> sin((size_t)out.n_sites);
            for (int i = 0; i < out.n_sites; ++i) sin[(size_t)i] = where_(p.users[(size_t)i]);
            vgl_gvcf_out g;
            check(vgl_gvcf_merge(ctx_of(slot), slot_of(slot), sin.data(), gvcf_dps_.data(), (int32_t)gvcf_dps_.size(), &g), "vgl_gvcf_merge");
            cur_out_ = &out;
            cur_pending_ = &p;
            stitch_->feed(g, sin.data(), out.n_sites, [&](GvcfStitcher::Block& b, int32_t site) {
                const vgl_site_out& so = out.sites[site];
                b.alleles = alleles_string(so, prm_.do_unobserved, prm_.do_gvcf);
                b.qs.assign(so.qs, so.qs + so.n_alleles);
                b.user = p.users[(size_t)site];
            });
            for (int i = 0; i < out.n_sites; ++i) // skipped sites are reported as before
                if (out.sites[i].skip_code != 0) deliver(out, p, i);
        } else {
            for (int i = 0; i < out.n_sites; ++i) deliver(out, p, i);
        }
        p.in_flight = false;
    }

    // one site's results -> the record callback
    void deliver(const vgl_batch_out& out, const Pending& p, int i)
    {
        SimRecordView v;
        v.nSamples = out.n_samples;
        const int S = out.n_samples;
        {
            const vgl_site_out& s = out.sites[i];
            v.site_id = p.first + i;
            v.user = p.users[(size_t)i];
            v.ret = s.skip_code;
            v.nAlleles = s.n_alleles;
            v.nAllelesObserved = s.n_alleles_observed;
            v.nGenotypes = s.n_genotypes;
            v.allele_unobserved = -1;
            for (int a = 0; a < 5; ++a) {
                v.alleles2acgt[a] = s.alleles2acgt[a];
                v.acgt2alleles[a] = s.acgt2alleles[a];
                if (s.alleles2acgt[a] == 4) v.allele_unobserved = a;
            }
            v.alleles = s.skip_code == 0 ? alleles_string(s, prm_.do_unobserved, prm_.do_gvcf) : std::string();
            v.info_dp_arr[0] = s.info_dp;
            v.gl_arr = out.gl ? out.gl + s.g_off : nullptr;
            v.gp_arr = out.gp ? out.gp + s.g_off : nullptr;
            if (out.narrow_bits) {
                const size_t c0 = (size_t)i * S, nG = (size_t)S * s.n_genotypes, nR = (size_t)S * s.n_alleles;
                v.fmt_dp_arr = widen(w_dp_, out.dp_n, out.narrow_bits, c0, (size_t)S);
                v.pl_arr = nullptr;
                if (out.pl_u8 && s.skip_code == 0) { // a cell without reads has a missing PL (vgl.h)
                    w_pl_.resize(nG);
                    const uint8_t* src = out.pl_u8 + s.g_off;
                    for (int smp = 0; smp < S; ++smp)
                        for (int g = 0; g < s.n_genotypes; ++g)
                            w_pl_[(size_t)smp * s.n_genotypes + g] = v.fmt_dp_arr[smp] ? (int32_t)src[(size_t)smp * s.n_genotypes + g] : VGL_I32_MISSING;
                    v.pl_arr = w_pl_.data();
                }
                v.fmt_ad_arr = out.ad_n && s.skip_code == 0 ? widen(w_ad_, out.ad_n, out.narrow_bits, (size_t)s.r_off, nR) : nullptr;
                v.fmt_adf_arr = out.adf_n && s.skip_code == 0 ? widen(w_adf_, out.adf_n, out.narrow_bits, (size_t)s.r_off, nR) : nullptr;
                v.fmt_adr_arr = out.adr_n && s.skip_code == 0 ? widen(w_adr_, out.adr_n, out.narrow_bits, (size_t)s.r_off, nR) : nullptr;
            } else {
                v.fmt_dp_arr = out.dp + (size_t)i * S;
                v.pl_arr = out.pl ? out.pl + s.g_off : nullptr;
                v.fmt_ad_arr = out.ad ? out.ad + s.r_off : nullptr;
                v.fmt_adf_arr = out.adf ? out.adf + s.r_off : nullptr;
                v.fmt_adr_arr = out.adr ? out.adr + s.r_off : nullptr;
            }
            v.info_ad_arr = s.info_ad;
            v.info_adf_arr = s.info_adf;
            v.info_adr_arr = s.info_adr;
            v.qs_arr = s.qs;
            v.i16_arr = s.i16;
            cb_(v);
        }
    }

    static const int32_t* widen(std::vector<int32_t>& dst, const void* src, int bits, size_t off, size_t n)
    {
        dst.resize(n);
        if (bits == 8) { const uint8_t* p = (const uint8_t*)src + off; for (size_t k = 0; k < n; ++k) dst[k] = p[k]; }
        else { const uint16_t* p = (const uint16_t*)src + off; for (size_t k = 0; k < n; ++k) dst[k] = p[k]; }
        return dst.data();
    }

    std::vector<int32_t> w_dp_, w_pl_, w_ad_, w_adf_, w_adr_;
    vgl_params prm_;
    Callback cb_;
    vgl_ctx* ctx_ = nullptr;      // = ctxs_[0]
    std::vector<vgl_ctx*> ctxs_;  // one per device
    int n_lanes_ = 0;             // devices x slots
    std::vector<Pending> pending_;
    std::unique_ptr<GvcfStitcher> stitch_;
    std::vector<int32_t> gvcf_dps_;
    std::function<vgl_gvcf_site_in(void*)> where_;
    const vgl_batch_out* cur_out_ = nullptr;
    const Pending* cur_pending_ = nullptr;
    uint8_t* in_ = nullptr;
    int cur_ = 0;
    int32_t fill_ = 0;
    int64_t next_site_ = 0;
};

// the multi-GPU driver is the batch simulator with a device list (see its second constructor)
using MultiGpuSimulator = BatchSimulator;

// ---------------------------------------------------------------------------------------------------------------
// VGL_HOST_BCF: the reference's whole write path for a site -- add_tags() (bcf_utils.cpp:426-507) and bcf_write()
// (htslib/vcf.c:1951-2001) -- happens on the device; the host only forwards what the input record passes through
// and appends the returned bytes to the (uncompressed or BGZF) output stream.
//
//   reference                                            here
//   ---------                                            ----
//   simulate_record_values(sim) + write_record_values()  BcfStreamSimulator::push_site(gts, in_rec fields)
//   bcf_write(out_fp, hdr, rec) per record               on_records(bytes, n_bytes, n_sites, n_skipped) per batch
class BcfStreamSimulator {
public:
    using Sink = std::function<void(const uint8_t* records, size_t n_bytes, int n_sites, int n_skipped)>;

    // dict: bcf_hdr_id2int(out_hdr, BCF_DT_ID, "DP" / "GL" / ...) of the output header
    // bgzf: the sink receives the records as BGZF blocks compressed on the device (VGL_HOST_BGZF: the reference's -O b; the host
    //       writes them behind its own header block(s) with hwrite and ends the file with the BGZF EOF block); not with enable_gvcf
    BcfStreamSimulator(vgl_params params, const vgl_bcf_dict& dict, Sink on_records, bool bgzf = false)
        : prm_(params), sink_(std::move(on_records)), bgzf_(bgzf)
    {
        prm_.abi_version = VGL_ABI_VERSION;
        prm_.host_output = bgzf ? VGL_HOST_BGZF : VGL_HOST_BCF;
        prm_.bcf_dict = dict;
        if (prm_.n_slots < 2) prm_.n_slots = 2;
        if (prm_.bcf_blob_bytes_per_site == 0) prm_.bcf_blob_bytes_per_site = 16;
        const int rc = vgl_create(&prm_, &ctx_);
        if (rc != VGL_OK) throw Error(rc, std::string("vgl_create: ") + vgl_strerror(rc));
        in_flight_.assign(prm_.n_slots, false);
        open_slot();
    }
    ~BcfStreamSimulator() { vgl_destroy(ctx_); }
    BcfStreamSimulator(const BcfStreamSimulator&) = delete;
    BcfStreamSimulator& operator=(const BcfStreamSimulator&) = delete;

    // One input record (after check_rec_alleles, vcfgl.cpp:75-163).  id / flt_info: with in_rec unpacked,
    //   id       = in_rec->shared.s[0 .. unpack_size[0])                              (nullptr / 0: ".")
    //   flt_info = in_rec->shared.s[unpack_size[0] + unpack_size[1] .. shared.l)      (nullptr / 0: FILTER ".", no INFO)
    //   fmt      = the FORMAT blocks of the input record besides GT, in its order: in_rec->indiv.s with the GT block taken out
    //              (typed key, descriptor, n_samples vectors each), n_fmt of them                (nullptr / 0: FORMAT is GT alone)
    void push_site(const int* true_gts_acgt_int, int32_t rid, int32_t pos, float qual, const uint8_t* id, uint32_t id_len,
                   const uint8_t* flt_info, uint32_t flt_info_len, uint32_t n_info, const uint8_t* fmt = nullptr, uint32_t fmt_len = 0,
                   uint32_t n_fmt = 0)
    {
        if (blob_fill_ + id_len + flt_info_len + fmt_len > (size_t)blob_cap_) {
            if (fill_ == 0) throw Error(VGL_EINVAL, "pass-through fields of one record exceed the blob (raise bcf_blob_bytes_per_site)");
            submit_current();
        }
        uint8_t* row = gt_ + (size_t)fill_ * prm_.n_samples;
        for (int s = 0; s < prm_.n_samples; ++s) {
            const int h0 = true_gts_acgt_int[2 * s], h1 = true_gts_acgt_int[2 * s + 1];
            row[s] = VGL_GT_PACK(h0 < 0 ? VGL_GT_MISSING : h0, h1 < 0 ? VGL_GT_MISSING : h1);
        }
        vgl_bcf_site_in& r = sin_[fill_];
        r.rid = rid;
        r.pos = pos;
        memcpy(&r.qual_bits, &qual, 4);
        r.n_info = n_info;
        r.id_off = (uint32_t)blob_fill_;
        r.id_len = id_len;
        if (id_len) memcpy(blob_ + blob_fill_, id, id_len);
        blob_fill_ += id_len;
        r.flt_info_off = (uint32_t)blob_fill_;
        r.flt_info_len = flt_info_len;
        if (flt_info_len) memcpy(blob_ + blob_fill_, flt_info, flt_info_len);
        blob_fill_ += flt_info_len;
        r.fmt_off = (uint32_t)blob_fill_;
        r.fmt_len = fmt_len;
        r.n_fmt = n_fmt;
        r._pad = 0;
        if (fmt_len) memcpy(blob_ + blob_fill_, fmt, fmt_len);
        blob_fill_ += fmt_len;
        if (++fill_ == prm_.max_batch_sites) submit_current();
    }

    // -doGVCF 1: the block merger runs on the device too (vgl_set_gvcf_dps); the stream then holds regular and block records
    // in output order, the seam between batches stitched by vgl_wait; the last open block arrives with finish()
    void enable_gvcf(const std::vector<int32_t>& gvcf_dps)
    {
        if (bgzf_) throw Error(VGL_EINVAL, "gVCF blocks are stitched across batches on the uncompressed stream: use VGL_HOST_BCF");
        check(vgl_set_gvcf_dps(ctx_, gvcf_dps.data(), (int32_t)gvcf_dps.size()), "vgl_set_gvcf_dps");
        gvcf_ = true;
    }

    void finish()
    {
        if (fill_ > 0) submit_current();
        for (int k = 0; k < prm_.n_slots; ++k) drain((cur_ + k) % prm_.n_slots);
        if (gvcf_) {
            const uint8_t* rec = nullptr;
            int64_t nb = 0;
            check(vgl_gvcf_flush(ctx_, &rec, &nb), "vgl_gvcf_flush");
            if (nb > 0) sink_(rec, (size_t)nb, 0, 0);
        }
    }

private:
    void check(int rc, const char* what)
    {
        if (rc != VGL_OK) throw Error(rc, std::string(what) + ": " + vgl_strerror(rc) + " (" + vgl_last_error(ctx_) + ")");
    }
    void open_slot()
    {
        drain(cur_);
        int64_t cap = 0;
        check(vgl_input_buffer(ctx_, cur_, &gt_, &cap), "vgl_input_buffer");
        check(vgl_bcf_input_buffer(ctx_, cur_, &sin_, &blob_, &blob_cap_), "vgl_bcf_input_buffer");
        fill_ = 0;
        blob_fill_ = 0;
    }
    void submit_current()
    {
        check(vgl_submit(ctx_, cur_, next_site_, fill_, nullptr, 0), "vgl_submit");
        in_flight_[cur_] = true;
        next_site_ += fill_;
        cur_ = (cur_ + 1) % prm_.n_slots;
        open_slot();
    }
    void drain(int slot)
    {
        if (!in_flight_[slot]) return;
        vgl_batch_out out;
        check(vgl_wait(ctx_, slot, &out), "vgl_wait");
        if (out.status != VGL_OK) throw Error(out.status, vgl_strerror(out.status));
        int skipped = 0;
        for (int i = 0; i < out.n_sites; ++i) skipped += out.sites[i].skip_code != 0; // nSitesSkipped, vcfgl.cpp:1553-1558
        if (bgzf_) sink_(out.bgzf, (size_t)out.bgzf_bytes, out.n_sites, skipped);
        else sink_(out.bcf, (size_t)out.bcf_bytes, out.n_sites, skipped);
        in_flight_[slot] = false;
    }

    vgl_params prm_;
    Sink sink_;
    vgl_ctx* ctx_ = nullptr;
    std::vector<bool> in_flight_;
    uint8_t *gt_ = nullptr, *blob_ = nullptr;
    bool gvcf_ = false, bgzf_ = false;
    vgl_bcf_site_in* sin_ = nullptr;
    int64_t blob_cap_ = 0;
    size_t blob_fill_ = 0;
    int cur_ = 0;
    int32_t fill_ = 0;
    int64_t next_site_ = 0;
};

// ---------------------------------------------------------------------------------------------------------------
// Input path: VCF text in, one callback per simulated site out -- the reference's main_simulate_record_values()
// (vcfgl.cpp:1469-1620) with the per-record work (vcf_parse, bcf_get_genotypes, check_rec_alleles) done on the device.
//
//   reference                                            here
//   ---------                                            ----
//   bcf_hdr_read: samples, ##contig lengths              VcfTextSimulator::read_header()
//   bcf_read + check_rec_alleles per record              vgl_parse_vcf() per chunk of text (k_vcf_lines, k_vcf_gt)
//   -explode 1 blank records (vcfgl.cpp:1489-1538)       row_map entries < 0 (vgl_place_rows fills REF|REF)
//   skip codes -1 / -2 (--rm-invar-sites 1|2)            vgl_in_site::skip_code, such records never become sites
//   ERROR()/ASSERT() on a malformed record               vgl::Error(VGL_EINVAL) naming the record and the vgl_in_status
class VcfTextSimulator {
public:
    struct Site { // what the callback gets besides the tags
        std::string contig;
        int64_t pos = 0;     // 0-based
        int64_t record = -1; // >= 0: n-th record of the input; -1: -explode site
        int32_t rid = 0;     // running number of the contig (changes when CHROM changes)
    };
    using Callback = std::function<void(const SimRecordView&, const Site&)>;

    // params.n_samples may be 0: it is taken from the #CHROM line
    VcfTextSimulator(vgl_params params, int gt_source, int explode, Callback cb)
        : prm_(params), source_(gt_source), explode_(explode), cb_(std::move(cb))
    {
    }
    ~VcfTextSimulator()
    {
        if (ps_) vgl_parser_destroy(ps_);
    }
    VcfTextSimulator(const VcfTextSimulator&) = delete;
    VcfTextSimulator& operator=(const VcfTextSimulator&) = delete;

    // -doGVCF 1: blocks through on_block, regular records through the constructor's callback (BatchSimulator::enable_gvcf)
    void enable_gvcf(std::vector<int32_t> gvcf_dps, std::function<void(const GvcfStitcher::Block&)> on_block)
    {
        gvcf_dps_ = std::move(gvcf_dps);
        on_block_ = std::move(on_block);
    }

    int64_t n_sites() const { return n_sites_; }
    int64_t n_skipped_input() const { return n_skipped_; }
    const std::vector<std::string>& samples() const { return samples_; }

    void run(FILE* in)
    {
        const std::string carry = read_header(in);
        if (prm_.n_samples == 0) prm_.n_samples = (int32_t)samples_.size();
        if ((size_t)prm_.n_samples != samples_.size()) throw Error(VGL_EINVAL, "n_samples does not match the #CHROM line");
        BatchSimulator sim(prm_, [&](const SimRecordView& v) { cb_(v, *static_cast<const Site*>(v.user)); });
        if (!gvcf_dps_.empty())
            sim.enable_gvcf(gvcf_dps_, [](void* u) { const Site* st = static_cast<const Site*>(u); return vgl_gvcf_site_in{st->rid, (int32_t)st->pos}; }, on_block_);
        const int32_t cap = sim.params().max_batch_sites;
        const int64_t text_cap = (int64_t)cap * (4 * (int64_t)prm_.n_samples + 64) + (1 << 16);
        int rc = vgl_parser_create(sim.context(), text_cap, cap, &ps_);
        if (rc != VGL_OK) throw Error(rc, std::string("vgl_parser_create: ") + vgl_last_error(sim.context()));
        uint8_t* text = nullptr;
        int64_t tcap = 0;
        vgl_parser_text_buffer(ps_, &text, &tcap);
        // Site records of the batches in flight: a batch's entries must outlive its delivery, which happens at the
        // latest when its slot comes round again (n_slots submissions later)
        ring_.assign((size_t)sim.params().n_slots + 1, std::vector<Site>());
        ring_at_ = 0;
        size_t have = carry.size();
        if ((int64_t)have > tcap) throw Error(VGL_EINVAL, "header tail larger than the text buffer");
        memcpy(text, carry.data(), have);
        bool eof = false;
        int64_t n_records_total = 0;
        int last_acgt0 = -1;
        for (;;) {
            if (!eof && (int64_t)have < tcap) {
                const size_t got = fread(text + have, 1, (size_t)tcap - have, in);
                have += got;
                if (got == 0) eof = true;
            }
            if (have == 0) break;
            vgl_parse_out po;
            rc = vgl_parse_vcf(ps_, (int64_t)have, source_, eof ? VGL_PARSE_FINAL : 0, &po);
            if (rc != VGL_OK) throw Error(rc, "vgl_parse_vcf failed");
            if (po.n_errors) {
                const vgl_in_site& b = po.sites[po.first_error_record];
                throw Error(VGL_EINVAL, "malformed record at position " + std::to_string(b.pos + 1) + " (vgl_in_status " +
                                            std::to_string(b.status) + "): " +
                                            std::string((const char*)text + b.line_off, std::min<size_t>(b.line_len, 80)));
            }
            if (po.n_records == 0) {
                if ((int64_t)have == tcap) throw Error(VGL_EINVAL, "a record does not fit the text buffer");
                if (eof) break;
                continue;
            }
            // the sites of this chunk (vcfgl.cpp:1479-1565), submitted batch by batch
            begin_batch(cap);
            for (int32_t i = 0; i < po.n_records; ++i) {
                const vgl_in_site& r = po.sites[i];
                const char* line = (const char*)text + r.line_off;
                const char* tab = (const char*)memchr(line, '\t', r.line_len);
                const size_t clen = tab ? (size_t)(tab - line) : r.line_len;
                if (contig_.size() != clen || memcmp(contig_.data(), line, clen) != 0) { // contig change (vcfgl.cpp:1484-1488)
                    contig_.assign(line, clen);
                    n_in_contig_ = 0;
                    ++rid_;
                }
                last_acgt0 = r.allele_acgt[0];
                if (explode_) {
                    if (r.pos < n_in_contig_) throw Error(VGL_EINVAL, "-explode 1 needs increasing positions within a contig");
                    if (r.pos != n_in_contig_ && fill_acgt_ < 0) fill_acgt_ = r.allele_acgt[0]; // explode_rec: blank copy of THIS record
                    for (; n_in_contig_ < r.pos; ++n_in_contig_) {
                        if (prm_.rm_invar_sites & 1) ++n_skipped_; // all hom-ref: check_rec_alleles returns -1
                        else add_site(sim, cap, n_in_contig_, -1, -1);
                    }
                }
                if (r.skip_code != 0) ++n_skipped_;
                else add_site(sim, cap, r.pos, i, n_records_total + i);
                ++n_in_contig_;
            }
            flush(sim, cap); // the next parse overwrites the rows
            n_records_total += po.n_records;
            const size_t used = (size_t)po.bytes_consumed;
            memmove(text, text + used, have - used);
            have -= used;
        }
        if (explode_ && !contig_.empty()) { // to the end of the LAST contig (vcfgl.cpp:1567-1611)
            const auto it = contigs_.find(contig_);
            const int64_t size = it == contigs_.end() ? 0 : it->second;
            if (fill_acgt_ < 0) fill_acgt_ = last_acgt0;
            if (prm_.rm_invar_sites & 1) {
                n_skipped_ += std::max<int64_t>(0, size - n_in_contig_);
            } else {
                begin_batch(cap);
                for (; n_in_contig_ < size; ++n_in_contig_) add_site(sim, cap, n_in_contig_, -1, -1);
                flush(sim, cap);
            }
        }
        sim.finish();
        vgl_parser_destroy(ps_);
        ps_ = nullptr;
    }

    // The same driver loop for an uncompressed BCF stream (what `bcftools view -Ou` or the reference's own -O u writes; a BGZF
    // stream must be inflated by the caller): header -> samples, contig names in rid order, the dictionary id of FORMAT/GT;
    // records -> vgl_parse_bcf() (k_bcf_gt) chunk by chunk.
    void run_bcf(FILE* in)
    {
        char magic[5];
        uint32_t l_text = 0;
        if (fread(magic, 1, 5, in) != 5 || memcmp(magic, "BCF\2\2", 5) != 0 || fread(&l_text, 4, 1, in) != 1) throw Error(VGL_EINVAL, "not an uncompressed BCF2 stream");
        std::string text(l_text, '\0');
        if (fread(&text[0], 1, l_text, in) != l_text) throw Error(VGL_EINVAL, "truncated BCF header");
        std::vector<std::string> rid_names;
        int gt_key = -1;
        {   // dictionary of FILTER / INFO / FORMAT ids: PASS = 0, then first appearance, IDX= overrides (htslib/vcf.c bcf_hdr_sync)
            std::map<std::string, int> ids;
            ids["PASS"] = 0;
            int next = 1;
            size_t p = 0;
            while (p < text.size()) {
                size_t e = text.find('\n', p);
                if (e == std::string::npos) e = text.size();
                const std::string line = text.substr(p, e - p);
                p = e + 1;
                const bool dict = line.rfind("##FILTER=<", 0) == 0 || line.rfind("##INFO=<", 0) == 0 || line.rfind("##FORMAT=<", 0) == 0;
                if (dict || line.rfind("##contig=<", 0) == 0) {
                    const size_t id = line.find("ID=");
                    if (id == std::string::npos) continue;
                    const std::string name = line.substr(id + 3, line.find_first_of(",>", id) - id - 3);
                    if (!dict) {
                        const size_t len = line.find("length=");
                        contigs_[name] = len == std::string::npos ? 0 : atoll(line.c_str() + len + 7);
                        rid_names.push_back(name);
                        continue;
                    }
                    const size_t idx = line.find("IDX=");
                    if (idx != std::string::npos) ids[name] = atoi(line.c_str() + idx + 4), next = std::max(next, ids[name] + 1);
                    else if (!ids.count(name)) ids[name] = next++;
                    if (line.rfind("##FORMAT=<", 0) == 0 && name == "GT") gt_key = ids[name];
                } else if (line.rfind("#CHROM", 0) == 0) {
                    size_t q = 0;
                    for (int col = 0; q != std::string::npos; ++col) {
                        const size_t t = line.find('\t', q);
                        if (col >= 9) samples_.push_back(line.substr(q, t == std::string::npos ? t : t - q));
                        q = t == std::string::npos ? t : t + 1;
                    }
                }
            }
        }
        if (samples_.empty()) throw Error(VGL_EINVAL, "the BCF has no samples");
        if (gt_key < 0) throw Error(VGL_EINVAL, "Could not find GT tag in the BCF header");
        if (prm_.n_samples == 0) prm_.n_samples = (int32_t)samples_.size();
        if ((size_t)prm_.n_samples != samples_.size()) throw Error(VGL_EINVAL, "n_samples does not match the BCF header");
        BatchSimulator sim(prm_, [&](const SimRecordView& v) { cb_(v, *static_cast<const Site*>(v.user)); });
        if (!gvcf_dps_.empty())
            sim.enable_gvcf(gvcf_dps_, [](void* u) { const Site* st = static_cast<const Site*>(u); return vgl_gvcf_site_in{st->rid, (int32_t)st->pos}; }, on_block_);
        const int32_t cap = sim.params().max_batch_sites;
        const int64_t text_cap = (int64_t)cap * (2 * (int64_t)prm_.n_samples + 256) + (1 << 16);
        int rc = vgl_parser_create(sim.context(), text_cap, cap, &ps_);
        if (rc != VGL_OK) throw Error(rc, std::string("vgl_parser_create: ") + vgl_last_error(sim.context()));
        uint8_t* buf = nullptr;
        int64_t tcap = 0;
        vgl_parser_text_buffer(ps_, &buf, &tcap);
        ring_.assign((size_t)sim.params().n_slots + 1, std::vector<Site>());
        ring_at_ = 0;
        size_t have = 0;
        bool eof = false;
        int64_t n_records_total = 0;
        int last_acgt0 = -1, last_rid = -1;
        std::vector<uint32_t> off;
        for (;;) {
            if (!eof && (int64_t)have < tcap) {
                const size_t got = fread(buf + have, 1, (size_t)tcap - have, in);
                have += got;
                if (got == 0) eof = true;
            }
            if (have == 0) break;
            off.assign(1, 0u);
            for (size_t o = 0; o + 8 <= have && (int32_t)off.size() <= cap;) { // hop l_shared + l_indiv + 8
                uint32_t ls, li;
                memcpy(&ls, buf + o, 4);
                memcpy(&li, buf + o + 4, 4);
                if (o + 8 + ls + li > have) break;
                o += 8 + (size_t)ls + li;
                off.push_back((uint32_t)o);
            }
            const int32_t n = (int32_t)off.size() - 1;
            if (n == 0) {
                if (eof || (int64_t)have == tcap) throw Error(VGL_EINVAL, "truncated BCF record (or one that does not fit the buffer)");
                continue;
            }
            vgl_parse_out po;
            rc = vgl_parse_bcf(ps_, (int64_t)off[(size_t)n], off.data(), n, source_, gt_key, 0, &po);
            if (rc != VGL_OK) throw Error(rc, "vgl_parse_bcf failed");
            if (po.n_errors) {
                const vgl_in_site& b = po.sites[po.first_error_record];
                throw Error(VGL_EINVAL, "malformed record at position " + std::to_string(b.pos + 1) + " (vgl_in_status " + std::to_string(b.status) + ")");
            }
            begin_batch(cap);
            for (int32_t i = 0; i < n; ++i) {
                const vgl_in_site& r = po.sites[i];
                int32_t rid;
                memcpy(&rid, buf + r.line_off + 8, 4);
                if (rid != last_rid) { // contig change (vcfgl.cpp:1484-1488)
                    if (rid < 0 || (size_t)rid >= rid_names.size()) throw Error(VGL_EINVAL, "record with a contig id that is not in the header");
                    contig_ = rid_names[(size_t)rid];
                    last_rid = rid;
                    n_in_contig_ = 0;
                    ++rid_;
                }
                last_acgt0 = r.allele_acgt[0];
                if (explode_) {
                    if (r.pos < n_in_contig_) throw Error(VGL_EINVAL, "-explode 1 needs increasing positions within a contig");
                    if (r.pos != n_in_contig_ && fill_acgt_ < 0) fill_acgt_ = r.allele_acgt[0];
                    for (; n_in_contig_ < r.pos; ++n_in_contig_) {
                        if (prm_.rm_invar_sites & 1) ++n_skipped_;
                        else add_site(sim, cap, n_in_contig_, -1, -1);
                    }
                }
                if (r.skip_code != 0) ++n_skipped_;
                else add_site(sim, cap, r.pos, i, n_records_total + i);
                ++n_in_contig_;
            }
            flush(sim, cap);
            n_records_total += n;
            const size_t used = off[(size_t)n];
            memmove(buf, buf + used, have - used);
            have -= used;
        }
        if (explode_ && !contig_.empty()) { // to the end of the LAST contig (vcfgl.cpp:1567-1611)
            const auto it = contigs_.find(contig_);
            const int64_t size = it == contigs_.end() ? 0 : it->second;
            if (fill_acgt_ < 0) fill_acgt_ = last_acgt0;
            if (prm_.rm_invar_sites & 1) n_skipped_ += std::max<int64_t>(0, size - n_in_contig_);
            else {
                begin_batch(cap);
                for (; n_in_contig_ < size; ++n_in_contig_) add_site(sim, cap, n_in_contig_, -1, -1);
                flush(sim, cap);
            }
        }
        sim.finish();
        vgl_parser_destroy(ps_);
        ps_ = nullptr;
    }

private:
    void begin_batch(int32_t cap)
    {
        map_.clear();
        ring_[ring_at_].clear();
        ring_[ring_at_].reserve((size_t)cap); // pointers into it are handed out: no reallocation afterwards
    }
    void add_site(BatchSimulator& sim, int32_t cap, int64_t pos, int32_t src, int64_t rec_no)
    {
        Site st;
        st.contig = contig_;
        st.pos = pos;
        st.record = rec_no;
        st.rid = rid_;
        ring_[ring_at_].push_back(std::move(st));
        map_.push_back(src);
        if ((int32_t)map_.size() == cap) flush(sim, cap);
    }
    void flush(BatchSimulator& sim, int32_t cap)
    {
        if (map_.empty()) return;
        std::vector<Site>& sites = ring_[ring_at_];
        std::vector<void*> users(map_.size());
        for (size_t k = 0; k < map_.size(); ++k) users[k] = &sites[k];
        const uint8_t fill = fill_acgt_ >= 0 ? (uint8_t)(fill_acgt_ * 0x11) : 0;
        sim.push_device_sites(ps_, map_.data(), (int32_t)map_.size(), fill, users.data());
        n_sites_ += (int64_t)map_.size();
        ring_at_ = (ring_at_ + 1) % ring_.size();
        begin_batch(cap);
    }

    // header lines up to #CHROM; returns the bytes already read past it
    std::string read_header(FILE* in)
    {
        std::string buf, line;
        std::vector<char> tmp(1 << 16);
        size_t off = 0;
        for (;;) {
            size_t nl;
            while ((nl = buf.find('\n', off)) == std::string::npos) {
                const size_t got = fread(tmp.data(), 1, tmp.size(), in);
                if (got == 0) throw Error(VGL_EINVAL, "no #CHROM line in the VCF header");
                buf.append(tmp.data(), got);
            }
            line.assign(buf, off, nl - off);
            if (!line.empty() && line.back() == '\r') line.pop_back();
            off = nl + 1;
            if (line.rfind("##contig=<", 0) == 0) {
                const size_t id = line.find("ID="), len = line.find("length=");
                if (id != std::string::npos) {
                    const size_t e = line.find_first_of(",>", id);
                    contigs_[line.substr(id + 3, e - id - 3)] = len == std::string::npos ? 0 : atoll(line.c_str() + len + 7);
                }
            } else if (line.rfind("#CHROM", 0) == 0) {
                size_t p = 0;
                for (int col = 0; p != std::string::npos; ++col) {
                    const size_t e = line.find('\t', p);
                    if (col >= 9) samples_.push_back(line.substr(p, e == std::string::npos ? e : e - p));
                    p = e == std::string::npos ? e : e + 1;
                }
                if (samples_.empty()) throw Error(VGL_EINVAL, "the VCF has no sample columns");
                return buf.substr(off);
            } else if (line.rfind("##", 0) != 0)
                throw Error(VGL_EINVAL, "record before the #CHROM line");
        }
    }

    vgl_params prm_;
    int source_, explode_;
    Callback cb_;
    vgl_parser* ps_ = nullptr;
    std::vector<std::string> samples_;
    std::map<std::string, int64_t> contigs_;
    std::string contig_;
    int64_t n_in_contig_ = 0, n_sites_ = 0, n_skipped_ = 0;
    int fill_acgt_ = -1;
    int32_t rid_ = -1;
    std::vector<int32_t> gvcf_dps_;
    std::function<void(const GvcfStitcher::Block&)> on_block_;
    std::vector<std::vector<Site>> ring_;
    size_t ring_at_ = 0;
    std::vector<int32_t> map_;
};

} // namespace vgl
