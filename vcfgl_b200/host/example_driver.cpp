// example_driver.cpp -- minimal host driver over vgl_host.hpp (no htslib): simulates hom-ref/het/hom-alt
// genotypes for a few sites and prints one VCF-like line per record, the way the reference's
// write path would after add_tags().  Build: see vcfgl_b200/host/Makefile.
#include "vgl_host.hpp"

#include <cstdio>
#include <cstdlib>
#include <string>

int main(int argc, char** argv)
{
    const int S = 4, n_sites = argc > 1 ? atoi(argv[1]) : 6;
    vgl_params p;
    memset(&p, 0, sizeof p);
    p.n_samples = S;
    p.seed = 42;
    p.depth_mode = VGL_DEPTH_POISSON;
    p.depth_mean = 4.0;
    p.error_rate = 0.01;
    p.gl_model = 1;
    p.gl1_theta = 0.83;
    p.adjust_by = 0.499;
    p.do_unobserved = 1;
    p.tag_mask = VGL_TAG_GL | VGL_TAG_PL | VGL_TAG_FMT_DP | VGL_TAG_FMT_AD | VGL_TAG_INFO_DP;
    if (getenv("VGL_GVCF_DPS")) p.do_gvcf = 1;
    p.i16_mapq = 20;
    p.max_batch_sites = 4;
    p.n_slots = 2;
    p.host_output = getenv("VGL_NARROW") ? VGL_HOST_NARROW : VGL_HOST_I32; // same records either way
    if (const char* vcf = getenv("VGL_VCF_IN")) { // input path: VCF text file -> device parser -> simulation, one line per site
        FILE* f = fopen(vcf, "rb");
        if (!f) { perror(vcf); return 1; }
        p.n_samples = 0; // from the #CHROM line
        p.max_batch_sites = getenv("VGL_BATCH") ? atoi(getenv("VGL_BATCH")) : 4;
        p.rm_invar_sites = getenv("VGL_RM_INVAR") ? atoi(getenv("VGL_RM_INVAR")) : 0;
        const int source = getenv("VGL_SOURCE") ? atoi(getenv("VGL_SOURCE")) : 0, explode = getenv("VGL_EXPLODE") ? atoi(getenv("VGL_EXPLODE")) : 0;
        try {
            vgl::VcfTextSimulator sim(p, source, explode, [&](const vgl::SimRecordView& r, const vgl::VcfTextSimulator::Site& st) {
                if (r.ret < 0) { printf("%s\t%ld\tskipped(%d)\n", st.contig.c_str(), (long)st.pos + 1, r.ret); return; }
                printf("%s\t%ld\t%s\tDP=%d\trec=%ld\tDP:AD", st.contig.c_str(), (long)st.pos + 1, r.alleles.c_str(), r.info_dp_arr[0], (long)st.record);
                for (int s = 0; s < r.nSamples; ++s) {
                    printf("\t%d:", r.fmt_dp_arr[s]);
                    for (int a = 0; a < r.nAlleles; ++a) printf("%s%d", a ? "," : "", r.fmt_ad_arr[s * r.nAlleles + a]);
                }
                printf("\n");
            });
            if (const char* dps = getenv("VGL_GVCF_DPS")) { // -doGVCF 1 --gvcf-dps a,b,c: blocks merged on the device
                std::vector<int32_t> v;
                for (const char* q = dps; *q;) {
                    v.push_back(atoi(q));
                    while (*q && *q != ',') ++q;
                    if (*q == ',') ++q;
                }
                sim.enable_gvcf(v, [&](const vgl::GvcfStitcher::Block& b) {
                    const vgl::VcfTextSimulator::Site* st = static_cast<const vgl::VcfTextSimulator::Site*>(b.user);
                    printf("%s\t%ld\t%s\tBLOCK\tEND=%ld\tMIN_DP=%d\tn=%d\tDP:PL", st->contig.c_str(), (long)b.start + 1, b.alleles.c_str(), (long)b.end + 1, b.min_dp,
                           b.n_members);
                    for (int s = 0; s < (int)b.dp.size(); ++s) printf("\t%d:%d,%d,%d", b.dp[s], b.pl[3 * s], b.pl[3 * s + 1], b.pl[3 * s + 2]);
                    printf("\n");
                });
            }
            if (getenv("VGL_INPUT_IS_BCF")) sim.run_bcf(f); // uncompressed BCF (-O u)
            else sim.run(f);
            fclose(f);
            fprintf(stderr, "sites simulated: %ld, input-side skips: %ld\n", (long)sim.n_sites(), (long)sim.n_skipped_input());
            return 0;
        } catch (const vgl::Error& e) {
            fprintf(stderr, "vgl error %d: %s\n", e.status, e.what());
            return e.status == VGL_ENODEV ? 3 : 1;
        }
    }
    const bool zout = getenv("VGL_BGZF_OUT") != nullptr;
    if (const char* path = zout ? getenv("VGL_BGZF_OUT") : getenv("VGL_BCF_OUT")) {
        // VGL_BCF_OUT: an uncompressed BCF file (-O u), records serialised on the device (VGL_HOST_BCF).
        // VGL_BGZF_OUT: the reference's default container (-O b) -- the header goes out as a stored BGZF block written here, the
        // records as the BGZF blocks the device compressed (VGL_HOST_BGZF), then the 28-byte EOF block (htslib/bgzf.c).
        try {
            if (getenv("VGL_BATCH")) p.max_batch_sites = atoi(getenv("VGL_BATCH"));
            std::string hdr = "##fileformat=VCFv4.2\n##FILTER=<ID=PASS,Description=\"All filters passed\",IDX=0>\n##contig=<ID=1,length=1000,IDX=0>\n";
            hdr += "##FORMAT=<ID=DP,Number=1,Type=Integer,Description=\"depth\",IDX=1>\n##INFO=<ID=DP,Number=1,Type=Integer,Description=\"depth\",IDX=1>\n";
            hdr += "##FORMAT=<ID=GL,Number=G,Type=Float,Description=\"GL\",IDX=2>\n##FORMAT=<ID=PL,Number=G,Type=Integer,Description=\"PL\",IDX=3>\n";
            hdr += "##FORMAT=<ID=AD,Number=R,Type=Integer,Description=\"AD\",IDX=4>\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT";
            for (int s = 0; s < S; ++s) hdr += "\tind" + std::to_string(s + 1);
            hdr += "\n";
            FILE* f = fopen(path, "wb");
            if (!f) { perror(path); return 1; }
            const uint32_t l_text = (uint32_t)hdr.size() + 1;
            std::string head("BCF\2\2", 5);
            head.append(reinterpret_cast<const char*>(&l_text), 4);
            head.append(hdr.c_str(), l_text);
            if (!zout) fwrite(head.data(), 1, head.size(), f);
            else { // one stored deflate block inside a BGZF block (header texts beyond 65280 bytes would need several)
                uint32_t crc = 0xFFFFFFFFu;
                for (unsigned char c : head) {
                    crc ^= c;
                    for (int k = 0; k < 8; ++k) crc = (crc & 1u) ? (crc >> 1) ^ 0xEDB88320u : crc >> 1;
                }
                crc = ~crc;
                const uint16_t len = (uint16_t)head.size(), nlen = (uint16_t)~len, bsize = (uint16_t)(18 + 5 + head.size() + 8 - 1);
                const uint32_t isize = (uint32_t)head.size();
                const unsigned char gz[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
                fwrite(gz, 1, 16, f);
                fwrite(&bsize, 2, 1, f);
                fputc(1, f); // BFINAL = 1, BTYPE = 00
                fwrite(&len, 2, 1, f);
                fwrite(&nlen, 2, 1, f);
                fwrite(head.data(), 1, head.size(), f);
                fwrite(&crc, 4, 1, f);
                fwrite(&isize, 4, 1, f);
            }
            vgl_bcf_dict dict;
            memset(&dict, 0, sizeof dict);
            dict.dp = 1; dict.gl = 2; dict.pl = 3; dict.ad = 4;
            long n_rec_bytes = 0;
            vgl::BcfStreamSimulator sim(p, dict, [&](const uint8_t* rec, size_t n, int n_sites_b, int n_skipped) {
                fwrite(rec, 1, n, f); // the whole batch in one write: bcf_write() per record is gone
                n_rec_bytes += (long)n;
                fprintf(stderr, "batch: %d sites, %d skipped, %zu bytes\n", n_sites_b, n_skipped, n);
            }, zout);
            std::vector<int> gts(2 * S);
            const uint8_t pass[2] = {0x11, 0x00}; // FILTER=PASS as the VCF parser encodes it
            for (int i = 0; i < n_sites; ++i) {
                for (int s = 0; s < S; ++s) {
                    gts[2 * s] = (i + s) % 3 == 2 ? 1 : 0;
                    gts[2 * s + 1] = (i + s) % 3 >= 1 ? 1 : 0;
                }
                float qual;
                const uint32_t qmiss = VGL_F32_MISSING_BITS;
                memcpy(&qual, &qmiss, 4);
                sim.push_site(gts.data(), 0, 10 * i + 1, qual, nullptr, 0, pass, 2, 0);
            }
            sim.finish();
            if (zout) {
                const unsigned char eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
                fwrite(eof, 1, 28, f);
            }
            fclose(f);
            printf("wrote %s: %d sites, %ld record bytes\n", path, n_sites, n_rec_bytes);
            return 0;
        } catch (const vgl::Error& e) {
            fprintf(stderr, "vgl error %d: %s\n", e.status, e.what());
            return e.status == VGL_ENODEV ? 3 : 1;
        }
    }
    try {
        // VGL_DEVICES=0,1,...: consecutive batches alternate between the listed GPUs (vgl::MultiGpuSimulator); same output
        std::vector<int> devs;
        if (const char* dl = getenv("VGL_DEVICES"))
            for (const char* q = dl; *q;) {
                devs.push_back(atoi(q));
                while (*q && *q != ',') ++q;
                if (*q == ',') ++q;
            }
        if (devs.empty()) devs.push_back(0);
        if (getenv("VGL_BATCH")) p.max_batch_sites = atoi(getenv("VGL_BATCH"));
        if (getenv("VGL_SAMPLES_TAGS")) p.tag_mask |= VGL_TAG_QS | VGL_TAG_I16;
        std::vector<int32_t> gdps;
        if (const char* dps = getenv("VGL_GVCF_DPS"))
            for (const char* q = dps; *q;) {
                gdps.push_back(atoi(q));
                while (*q && *q != ',') ++q;
                if (*q == ',') ++q;
            }
        vgl::MultiGpuSimulator sim(p, devs, [&](const vgl::SimRecordView& r) {
            if (r.ret < 0) { printf("site %ld skipped (%d)\n", (long)r.site_id, r.ret); return; }
            printf("1\t%ld\t.\t%s\tDP=%d\tDP:AD:PL", (long)r.site_id + 1, r.alleles.c_str(), r.info_dp_arr[0]);
            for (int s = 0; s < r.nSamples; ++s) {
                printf("\t%d:", r.fmt_dp_arr[s]);
                for (int a = 0; a < r.nAlleles; ++a) printf("%s%d", a ? "," : "", r.fmt_ad_arr[s * r.nAlleles + a]);
                printf(":");
                for (int g = 0; g < r.nGenotypes; ++g) {
                    const int v = r.pl_arr[s * r.nGenotypes + g];
                    if (v == VGL_I32_MISSING) printf("%s.", g ? "," : "");
                    else printf("%s%d", g ? "," : "", v);
                }
            }
            printf("\n");
        });
        std::vector<vgl_gvcf_site_in> where((size_t)n_sites);
        if (!gdps.empty()) { // -doGVCF 1: blocks merged on the device, stitched across batches (and devices) on the host
            for (int i = 0; i < n_sites; ++i) { where[(size_t)i].rid = 0; where[(size_t)i].pos = i + (i >= n_sites / 2 ? 3 : 0); } // one gap in the middle
            sim.enable_gvcf(gdps, [&](void* u) { return *static_cast<vgl_gvcf_site_in*>(u); }, [&](const vgl::GvcfStitcher::Block& b) {
                printf("1\t%ld\t%s\tBLOCK\tEND=%ld\tMIN_DP=%d\tn=%d\tDP:PL", (long)b.start + 1, b.alleles.c_str(), (long)b.end + 1, b.min_dp, b.n_members);
                for (int s = 0; s < (int)b.dp.size(); ++s) printf("\t%d:%d,%d,%d", b.dp[s], b.pl[3 * s], b.pl[3 * s + 1], b.pl[3 * s + 2]);
                printf("\n");
            });
        }
        std::vector<int> gts(2 * S);
        const bool invar = getenv("VGL_INVARIANT") != nullptr; // mostly hom-ref sites (gVCF blocks form)
        for (int i = 0; i < n_sites; ++i) {
            for (int s = 0; s < S; ++s) { // A = REF, C = ALT (binary source, vcfgl.cpp:103-128)
                gts[2 * s] = (!invar || i % 7 == 0) && (i + s) % 3 == 2 ? 1 : 0;
                gts[2 * s + 1] = (!invar || i % 7 == 0) && (i + s) % 3 >= 1 ? 1 : 0;
            }
            sim.push_site(gts.data(), gdps.empty() ? nullptr : &where[(size_t)i]);
        }
        sim.finish();
    } catch (const vgl::Error& e) {
        fprintf(stderr, "vgl error %d: %s\n", e.status, e.what());
        return e.status == VGL_ENODEV ? 3 : 1;
    }
    return 0;
}
