// example_driver.cpp -- minimal host driver over vgl_host.hpp (no htslib): simulates hom-ref/het/hom-alt
// genotypes for a few sites and prints one VCF-like line per record, the way the reference's
// write path would after add_tags().  Build: see vcfgl_b200/host/Makefile.
#include "vgl_host.hpp"

#include <cstdio>
#include <cstdlib>

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
    p.i16_mapq = 20;
    p.max_batch_sites = 4;
    p.n_slots = 2;
    p.host_output = getenv("VGL_NARROW") ? VGL_HOST_NARROW : VGL_HOST_I32; // same records either way
    try {
        vgl::BatchSimulator sim(p, [&](const vgl::SimRecordView& r) {
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
        std::vector<int> gts(2 * S);
        for (int i = 0; i < n_sites; ++i) {
            for (int s = 0; s < S; ++s) { // A = REF, C = ALT (binary source, vcfgl.cpp:103-128)
                gts[2 * s] = (i + s) % 3 == 2 ? 1 : 0;
                gts[2 * s + 1] = (i + s) % 3 >= 1 ? 1 : 0;
            }
            sim.push_site(gts.data());
        }
        sim.finish();
    } catch (const vgl::Error& e) {
        fprintf(stderr, "vgl error %d: %s\n", e.status, e.what());
        return e.status == VGL_ENODEV ? 3 : 1;
    }
    return 0;
}
