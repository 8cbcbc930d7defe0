// multi_pool_main.cpp — drives h264b2_multi_decode (csrc/host/h264_multi.cpp) against tests/mock_engine.cpp; built with -fsanitize=thread
// by tests/test_multi_pool_tsan.py.  usage: multi_pool_main threads replicas flags file.h264 [file2.h264 ...]
#include "h264_multi_b200.h"
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>
int main(int argc, char **argv) {
    if (argc < 5) return 2;
    const int threads = atoi(argv[1]), replicas = atoi(argv[2]), flags = atoi(argv[3]);
    std::vector<const char *> paths;
    for (int r = 0; r < replicas; r++) for (int i = 4; i < argc; i++) paths.push_back(argv[i]);
    std::vector<uint64_t> hashes(paths.size());
    H264B2MultiStats st;
    char err[512] = "";
    const int rc = h264b2_multi_decode(0, (int)paths.size(), paths.data(), threads, flags, hashes.data(), &st, err, sizeof err);
    if (rc) { fprintf(stderr, "multi_decode failed (%d): %s\n", rc, err); return 1; }
    printf("{\"pictures\": %lld, \"frames_out\": %lld, \"units\": %d, \"threads\": %d", (long long)st.pictures, (long long)st.frames_out, st.units, st.threads);
    printf(", \"hashes\": [");
    for (size_t i = 0; i < hashes.size(); i++) printf("%s%llu", i ? ", " : "", (unsigned long long)hashes[i]);
    printf("]}\n");
    return 0;
}
