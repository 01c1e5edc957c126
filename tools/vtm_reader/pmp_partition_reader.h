// Binary PartitionMat reader for the VTM side (opt-in: it changes the consumer).
//
// The patched VTM-10.0 parses <seq>_<comp>_QP<qp>_PartitionMat.txt with one getline + std::stoi per value
// (codec/vtm10.0-source-with-pmp-fast-alg/App/EncoderApp/EncAppCfg.cpp:4301-4398: per frame hor[R][C], ver[R][C],
// qt[R/2][C/2], dire[3][R][C], R = 16*(H>>6), C = 16*(W>>6), :4246-4249).  A 4K x 30-frame sequence is 160 M lines per
// QP.  The B200 path already holds those values as int8 vectors in exactly that order (pmp_assemble_frames), so it can
// write them raw: <seq>_<comp>_QP<qp>_PartitionMat.bin = a 32-byte header + F * (2RC + RC/4 + 3RC) int8 values.
//
// pmp_read_partition() fills the reader's own arrays from the .bin when it exists and falls back to the reference's
// text parse otherwise, so one call replaces the body of the per-component loops at EncAppCfg.cpp:4301-4398:
//
//     pmp_read_partition(partitionMatPath + seqNameLuma,   partitionFrameNum, partitionRow, partitionColumn, 0,
//                        partitionHorMat, partitionVerMat, qtDepthMat, directionMat);
//     pmp_read_partition(partitionMatPath + seqNameChroma, partitionFrameNum, partitionRow, partitionColumn, 1,
//                        partitionHorMat, partitionVerMat, qtDepthMat, directionMat);
//
// Header-only, C++11, no dependencies.  Written from scratch for this repository (the reference has no such reader).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

struct PmpBinHeader {            // little-endian, 32 bytes
    char magic[8];               // "PMPPART1"
    int32_t frames, rows, cols;  // rows = R (4x4 units), cols = C
    int32_t reserved[3];
};

// values per frame in file order: hor | ver | qt | dire (Map2Partition.py:401-412)
inline size_t pmp_values_per_frame(int R, int C) { return (size_t)2 * R * C + (size_t)(R / 2) * (C / 2) + (size_t)3 * R * C; }

// Scatter one frame's values (file order) into the VTM reader's arrays for component k (0 luma, 1 chroma).
template <typename T>
inline void pmp_scatter_frame(const T *v, int frm, int k, int R, int C, uint8_t ****hor, uint8_t ****ver, uint8_t ****qt, int8_t *****dire)
{
    for (int i = 0; i < R; i++)
        for (int j = 0; j < C; j++) hor[frm][k][i][j] = (uint8_t)*v++;
    for (int i = 0; i < R; i++)
        for (int j = 0; j < C; j++) ver[frm][k][i][j] = (uint8_t)*v++;
    for (int i = 0; i < R / 2; i++)
        for (int j = 0; j < C / 2; j++) qt[frm][k][i][j] = (uint8_t)*v++;
    for (int d = 0; d < 3; d++)
        for (int i = 0; i < R; i++)
            for (int j = 0; j < C; j++) dire[frm][k][d][i][j] = (int8_t)*v++;
}

// base = path without extension.  Returns 0 on success (binary), 1 on success (text fallback), negative on error.
inline int pmp_read_partition(const std::string &base, int frames, int R, int C, int k, uint8_t ****hor, uint8_t ****ver,
                              uint8_t ****qt, int8_t *****dire)
{
    const size_t per = pmp_values_per_frame(R, C);
    if (FILE *fp = std::fopen((base + ".bin").c_str(), "rb")) {
        PmpBinHeader h;
        if (std::fread(&h, sizeof h, 1, fp) != 1 || std::memcmp(h.magic, "PMPPART1", 8) != 0 || h.rows != R || h.cols != C ||
            h.frames < frames) {
            std::fclose(fp);
            std::fprintf(stderr, "pmp_read_partition: %s.bin does not match %d frames of %d x %d units\n", base.c_str(), frames, R, C);
            return -2;
        }
        std::vector<int8_t> buf(per);
        for (int f = 0; f < frames; f++) {
            if (std::fread(buf.data(), 1, per, fp) != per) { std::fclose(fp); return -3; }
            pmp_scatter_frame(buf.data(), f, k, R, C, hor, ver, qt, dire);
        }
        std::fclose(fp);
        return 0;
    }
    // the reference's format: one decimal integer per line
    std::ifstream in((base + ".txt").c_str());
    if (!in) { std::fprintf(stderr, "pmp_read_partition: cannot open %s.txt\n", base.c_str()); return -1; }
    std::vector<int> buf(per);
    std::string line;
    for (int f = 0; f < frames; f++) {
        for (size_t i = 0; i < per; i++) {
            if (!std::getline(in, line)) return -3;
            buf[i] = std::atoi(line.c_str());
        }
        pmp_scatter_frame(buf.data(), f, k, R, C, hor, ver, qt, dire);
    }
    return 1;
}
