// Test helper (CPU): decodes one image file with the front end's reader (visgeom_b200/host/image_io.hpp) and writes
// "cols rows\n" followed by the grey pixels to stdout ("0 0\n" for a file the reader rejects).
#include <cstdio>

#include "image_io.hpp"

int main(int argc, char **argv)
{
    if (argc != 2) return 2;
    const visgeom_b200::Mat8u img = visgeom_b200::image_io::imread_grey(argv[1]);
    std::printf("%d %d\n", img.cols, img.rows);
    if (!img.empty()) std::fwrite(img.data.data(), 1, img.data.size(), stdout);
    return 0;
}
