// Host build of the criterion arithmetic that the CUDA kernels of custom_d_fine_b200/csrc/loss.cu run
// (csrc/loss_math.cuh is plain C++): the same work-item functions driven by serial loops instead of thread grids, so
// that the CPU test-suite (tests/test_loss_math_cpu.py) can pin the kernel arithmetic — values AND gradients — against
// the torch restatement of the reference criterion on a box without a GPU.  Test infrastructure only.
#include <cstring>
#include <vector>
#include "../../custom_d_fine_b200/csrc/loss_math.cuh"

using namespace lossmath;

extern "C" int loss_host_desc_size() { return (int)sizeof(dfine_loss_desc); }

extern "C" int loss_host_run(const dfine_loss_desc* desc, float* out, const float* gout, float* dlogits, float* dpre_logits,
                             float* denc_logits, float* dboxes, float* dpre_boxes, float* denc_boxes, float* dcorners) {
    LossDesc d = *desc;
    const int L = d.L, H = L + 2;
    d.Qm = d.Q > d.n_dn ? d.Q : d.n_dn;
    std::vector<int> maps((size_t)(L + 4) * d.B * d.Qm, -1);
    int cnt[2] = {0, 0};
    d.maps = maps.data();
    d.cnt = cnt;
    for (long j = 0; j < d.ncols; ++j) {
        int map, b, q, t;
        if (!table_entry(d, j, &map, &b, &q, &t)) continue;
        maps[((size_t)map * d.B + b) * d.Qm + q] = t;
        if (map == map_go(d)) cnt[0]++;
        if (map == map_dn(d)) cnt[1]++;
    }
    std::vector<double> acc(6 * H + 6 * L, 0.0);
    std::vector<int> notsame(2 * L, 0);
    const int groups = d.n_dn > 0 ? 2 : 1;
    for (int g = 0; g < groups; ++g) {
        for (int h = 0; h < n_heads(d, g); ++h) {
            const HeadView v = head_view(d, g, h);
            for (int b = 0; b < d.B; ++b)
                for (int q = 0; q < v.nq; ++q) acc[acc_off(d, 0) + g * H + h] += (double)vfl_row(d, g, h, b, q, nullptr, 0.f);
            const long n = g == 0 ? d.go_cap : d.n_dn_entries;
            for (long e = 0; e < n; ++e) {
                float l1, gl;
                long row;
                if (box_entry(d, g, h, e, &l1, &gl, nullptr, 0.f, 0.f, &row)) {
                    acc[acc_off(d, 1) + g * H + h] += l1;
                    acc[acc_off(d, 2) + g * H + h] += gl;
                }
            }
        }
        const int nq = g == 0 ? d.Q : d.n_dn;
        for (int l = 0; l < L; ++l)
            for (int b = 0; b < d.B; ++b)
                for (int q = 0; q < nq; ++q)
                    for (int e = 0; e < 4; ++e) {
                        const LocalOut o = local_item(d, g, l, b, q, e, nullptr, 0.f, 0.f, 0.f);
                        acc[acc_off(d, 3) + g * L + l] += o.fgl;
                        acc[acc_off(d, o.matched ? 4 : 5) + g * L + l] += o.per;
                        if (!o.same) notsame[g * L + l] = 1;
                    }
    }
    finalize(d, acc.data(), notsame.data(), out);
    if (!gout) return 0;
    for (int g = 0; g < groups; ++g) {
        for (int h = 0; h < n_heads(d, g); ++h) {
            const HeadView v = head_view(d, g, h);
            float* lbase = h < L ? dlogits + (long)h * d.B * d.Qt * d.C : (h == L ? dpre_logits : denc_logits);
            float* bbase = h < L ? dboxes + (long)h * d.B * d.Qt * 4 : (h == L ? dpre_boxes : denc_boxes);
            const float sv = gout[g * H + h] / norm_vfl(d, g);
            for (int b = 0; b < d.B; ++b)
                for (int q = 0; q < v.nq; ++q) vfl_row(d, g, h, b, q, lbase + ((long)b * v.ldb + v.q0 + q) * d.C, sv);
            const long n = g == 0 ? d.go_cap : d.n_dn_entries;
            const float s1 = gout[2 * H + g * H + h] / norm_box(d, g), s2 = gout[4 * H + g * H + h] / norm_box(d, g);
            for (long e = 0; e < n; ++e) {
                float l1, gl, db[4];
                long row;
                if (box_entry(d, g, h, e, &l1, &gl, db, s1, s2, &row)) std::memcpy(bbase + row * 4, db, 16);
            }
        }
        const int nq = g == 0 ? d.Q : d.n_dn;
        for (int l = 0; l < L; ++l) {
            const float c_fgl = gout[6 * H + g * L + l] / norm_box(d, g);
            const float gd = gout[6 * H + 2 * L + g * L + l];
            const float c_pos = out[6 * H + 4 * L + g * L + l] * gd, c_neg = out[6 * H + 6 * L + g * L + l] * gd;
            for (int b = 0; b < d.B; ++b)
                for (int q = 0; q < nq; ++q)
                    for (int e = 0; e < 4; ++e) {
                        float grow[NB_MAX] = {0};
                        local_item(d, g, l, b, q, e, grow, c_fgl, c_pos, c_neg);
                        const int qrow = (g == 0 ? d.n_dn : 0) + q;
                        std::memcpy(dcorners + ((((long)l * d.B + b) * d.Qt + qrow) * 4 + e) * d.NB, grow, sizeof(float) * d.NB);
                    }
        }
    }
    return 0;
}
