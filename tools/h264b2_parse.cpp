// h264b2_parse — run the native host front end (NAL split, CAVLC/CABAC, derivations) over an Annex-B stream.
//   h264b2_parse in.h264 out.bin [max_pictures]   write the per-picture structure-of-arrays as a picture container
//                                                 (the format CH264VideoDecoderB200::open and bench.py also read)
//   h264b2_parse --bench in.h264 [repeats]        parse only, report pictures/s of the host stage on one thread
#include "h264_front_b200.h"
#include <chrono>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
int main(int argc, char **argv) {
    if (argc >= 3 && !strcmp(argv[1], "--bench")) {
        const int reps = argc > 3 ? atoi(argv[3]) : 1;
        long pics = 0; size_t bytes = 0;
        std::vector<double> best;          // per picture: the fastest of the repeats (robust against a noisy host)
        const auto t0 = std::chrono::steady_clock::now();
        for (int r = 0; r < reps; r++) {
            H264B2Front *f = nullptr;
            if (h264b2_front_create(&f, nullptr, nullptr, nullptr) || h264b2_front_open_file(f, argv[2])) { fprintf(stderr, "cannot open %s\n", argv[2]); return 1; }
            size_t idx = 0;
            auto tp = std::chrono::steady_clock::now();
            for (;;) {
                H264B2FrontEvent ev;
                if (h264b2_front_next(f, &ev) < 0) { fprintf(stderr, "error: %s\n", h264b2_front_last_error(f)); return 1; }
                if (ev.kind == H264B2_EV_END) break;
                if (ev.kind == H264B2_EV_PICTURE) {
                    pics++; bytes += ev.block_bytes; h264b2_front_release(f, ev.block);
                    const auto tn = std::chrono::steady_clock::now();
                    const double d = std::chrono::duration<double>(tn - tp).count();
                    tp = tn;
                    if (idx >= best.size()) best.push_back(d); else if (d < best[idx]) best[idx] = d;
                    idx++;
                }
            }
            h264b2_front_destroy(f);
        }
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        double bs = 0; for (double d : best) bs += d;
        printf("{\"pictures\": %ld, \"seconds\": %.3f, \"pictures_per_s\": %.1f, \"best_pictures_per_s\": %.1f, \"soa_bytes_per_picture\": %.0f}\n", pics, s, pics / s,
               bs > 0 ? best.size() / bs : 0.0, pics ? (double)bytes / pics : 0.0);
        return 0;
    }
    if (argc < 3) { fprintf(stderr, "usage: %s in.h264 out.bin [max_pictures] | --bench in.h264 [repeats]\n", argv[0]); return 2; }
    return h264b2_front_write_container(argv[1], argv[2], argc > 3 ? atoi(argv[3]) : 0) ? 1 : 0;
}
