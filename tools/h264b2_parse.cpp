// h264b2_parse — run the native host front end (NAL split, CAVLC/CABAC, derivations) over an Annex-B stream and write the
// per-picture structure-of-arrays as a picture container (the format CH264VideoDecoderB200::open and bench.py read).
//   h264b2_parse in.h264 out.bin [max_pictures]
#include "h264_front_b200.h"
#include <stdio.h>
#include <stdlib.h>
int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s in.h264 out.bin [max_pictures]\n", argv[0]); return 2; }
    return h264b2_front_write_container(argv[1], argv[2], argc > 3 ? atoi(argv[3]) : 0) ? 1 : 0;
}
