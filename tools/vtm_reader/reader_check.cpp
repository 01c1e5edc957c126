// Test harness of pmp_partition_reader.h: reads <base> (binary if <base>.bin exists, else text) into arrays allocated the
// way EncAppCfg.cpp:4271-4298 allocates them and prints an FNV-1a checksum over every array in a fixed order.
//   reader_check <base> <frames> <R> <C>
#include "pmp_partition_reader.h"

int main(int argc, char **argv)
{
    if (argc < 5) return 2;
    const std::string base = argv[1];
    const int F = std::atoi(argv[2]), R = std::atoi(argv[3]), C = std::atoi(argv[4]);
    uint8_t ****hor = new uint8_t ***[F], ****ver = new uint8_t ***[F], ****qt = new uint8_t ***[F];
    int8_t *****dire = new int8_t ****[F];
    for (int f = 0; f < F; f++) {
        hor[f] = new uint8_t **[2]; ver[f] = new uint8_t **[2]; qt[f] = new uint8_t **[2]; dire[f] = new int8_t ***[2];
        for (int k = 0; k < 2; k++) {
            hor[f][k] = new uint8_t *[R]; ver[f][k] = new uint8_t *[R]; qt[f][k] = new uint8_t *[R >> 1]; dire[f][k] = new int8_t **[3];
            for (int i = 0; i < R; i++) { hor[f][k][i] = new uint8_t[C](); ver[f][k][i] = new uint8_t[C](); }
            for (int i = 0; i < (R >> 1); i++) qt[f][k][i] = new uint8_t[C >> 1]();
            for (int d = 0; d < 3; d++) {
                dire[f][k][d] = new int8_t *[R];
                for (int j = 0; j < R; j++) dire[f][k][d][j] = new int8_t[C]();
            }
        }
    }
    const int rc = pmp_read_partition(base, F, R, C, 0, hor, ver, qt, dire);
    if (rc < 0) return 1;
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](int v) { h = (h ^ (uint64_t)(uint8_t)v) * 1099511628211ull; };
    for (int f = 0; f < F; f++) {
        for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) mix(hor[f][0][i][j]);
        for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) mix(ver[f][0][i][j]);
        for (int i = 0; i < R / 2; i++) for (int j = 0; j < C / 2; j++) mix(qt[f][0][i][j]);
        for (int d = 0; d < 3; d++) for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) mix(dire[f][0][d][i][j]);
    }
    std::printf("%s %016llx\n", rc == 0 ? "bin" : "txt", (unsigned long long)h);
    return 0;
}
