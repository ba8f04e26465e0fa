// pfnav_pfmap.cu -- PFMAP terrain ingestion: the engine's ASCII map format straight into the
// device cost pass. Host code only; the cost grids themselves are made by k_cost_from_tiles (pfnav_fields.cu).
//
// Reference: docs/pfmap.txt; al_parse_pfmap_header (asset_load.c:168), M_AL_InitMapFromStream
// (map/map_asset_load.c:615: materials, splats, then rows * cols chunks of 32 x 32 tiles), m_al_read_row /
// m_al_read_pfchunk (:156-193: a line holds any number of 24-character tiles separated by blanks, a chunk is
// read line by line until 1024 tiles are in) and m_al_parse_tile (:103: fixed character positions).
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>
#include "pfnav_internal.cuh"

namespace {

#define PFMAP_MAX_LINE 256          /* MAX_LINE_LEN (asset_load.h:45): longer lines are cut there */

struct line_reader {
    const char *p, *end;
    // AL_ReadLine: one line without its terminator, cut to MAX_LINE_LEN - 1 characters; false at the end of the text
    bool next(std::string &out)
    {
        if (p >= end) return false;
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        const char *stop = nl ? nl : end;
        out.assign(p, (size_t)(stop - p));
        if (!out.empty() && out.back() == '\r') out.pop_back();
        if (out.size() > PFMAP_MAX_LINE - 1) out.resize(PFMAP_MAX_LINE - 1);
        p = nl ? nl + 1 : end;
        return true;
    }
};

inline int a2i(char c) { return c - '0'; }

// m_al_parse_tile (map_asset_load.c:103)
bool parse_tile(const char *s, size_t len, pfnav_tile *out)
{
    if (len != 24) return false;
    const char hex[2] = {s[0], '\0'};
    out->type = (int32_t)strtol(hex, nullptr, 16);
    out->base_height = (s[1] == '-' ? -1 : 1) * (10 * a2i(s[2]) + a2i(s[3]));
    out->ramp_height = 10 * a2i(s[4]) + a2i(s[5]);
    out->pathable = a2i(s[12]) != 0;
    return true;
}

}   // namespace

extern "C" int pfnav_pfmap_parse(const char *text, size_t len, int *out_chunk_rows, int *out_chunk_cols, pfnav_tile *out_tiles,
                                 size_t cap_tiles)
{
    PF_ARG(text && out_chunk_rows && out_chunk_cols, "null argument");
    line_reader rd = { text, text + len };
    std::string line;
    float version = 0.0f;
    int num_materials = 0, num_splats = 0, rows = 0, cols = 0;
    // al_parse_pfmap_header (asset_load.c:168)
    PF_ARG(rd.next(line) && sscanf(line.c_str(), "version %f", &version) == 1, "PFMAP: version line");
    PF_ARG(rd.next(line) && sscanf(line.c_str(), "num_materials %d", &num_materials) == 1, "PFMAP: num_materials line");
    if (version >= 1.1f)
        PF_ARG(rd.next(line) && sscanf(line.c_str(), "num_splats %d", &num_splats) == 1, "PFMAP: num_splats line");
    PF_ARG(rd.next(line) && sscanf(line.c_str(), "num_rows %d", &rows) == 1, "PFMAP: num_rows line");
    PF_ARG(rd.next(line) && sscanf(line.c_str(), "num_cols %d", &cols) == 1, "PFMAP: num_cols line");
    PF_ARG(num_materials >= 0 && num_splats >= 0 && rows > 0 && cols > 0 && rows <= 256 && cols <= 256, "PFMAP: header values");
    *out_chunk_rows = rows; *out_chunk_cols = cols;
    if (!out_tiles) return PFNAV_OK;
    const size_t ntiles = (size_t)rows * cols * 1024;
    PF_ARG(cap_tiles >= ntiles, "PFMAP: tile buffer too small");
    for (int i = 0; i < num_materials; i++)          // m_al_read_material (:195)
        PF_ARG(rd.next(line) && line.compare(0, 8, "material") == 0, "PFMAP: material line");
    for (int i = 0; i < num_splats; i++)             // m_al_read_splat (:216)
        PF_ARG(rd.next(line) && line.compare(0, 5, "splat") == 0, "PFMAP: splat line");
    for (size_t chunk = 0; chunk < (size_t)rows * cols; chunk++) {
        size_t have = 0;
        while (have < 1024) {                        // m_al_read_pfchunk (:181)
            PF_ARG(rd.next(line), "PFMAP: file ends inside the tile list");
            size_t pos = 0;
            while (true) {                           // m_al_read_row (:156): tokens separated by " \t\n"
                while (pos < line.size() && (line[pos] == ' ' || line[pos] == '\t')) pos++;
                if (pos >= line.size()) break;
                size_t stop = pos;
                while (stop < line.size() && line[stop] != ' ' && line[stop] != '\t') stop++;
                PF_ARG(have < 1024, "PFMAP: a tile row runs past the end of its chunk");
                PF_ARG(parse_tile(line.data() + pos, stop - pos, &out_tiles[chunk * 1024 + have]), "PFMAP: malformed tile (24 characters expected)");
                have++;
                pos = stop;
            }
        }
    }
    return PFNAV_OK;
}

// M_AL_InitMapFromStream + N_NewCtxForMapData (nav.c:2284) for the chosen layers: parse, create the map, make every
// layer's cost grid on the device from the tiles (layer i follows the reference layer ref_layers[i]) and build its
// local islands and portals.
extern "C" int pfnav_map_load_pfmap(pfnav_ctx *ctx, const char *text, size_t len, int nlayers, const int32_t *ref_layers,
                                    float map_x, float map_z)
{
    PF_ARG(ctx && text && ref_layers && nlayers > 0, "null argument / nlayers");
    PF_NEED_DEVICE(ctx);
    int rows = 0, cols = 0;
    int rc = pfnav_pfmap_parse(text, len, &rows, &cols, nullptr, 0);
    if (rc) return rc;
    std::vector<pfnav_tile> tiles((size_t)rows * cols * 1024);
    rc = pfnav_pfmap_parse(text, len, &rows, &cols, tiles.data(), tiles.size());
    if (rc) return rc;
    rc = pfnav_map_create(ctx, cols, rows, nlayers, map_x, map_z);
    if (rc) return rc;
    std::vector<const void *> ptrs((size_t)rows * cols);
    for (size_t i = 0; i < ptrs.size(); i++) ptrs[i] = tiles.data() + i * 1024;
    for (int l = 0; l < nlayers; l++) {
        rc = pfnav_map_cost_from_tiles(ctx, l, ref_layers[l], ptrs.data(), sizeof(pfnav_tile));
        if (rc) return rc;
        rc = pfnav_map_build_nav(ctx, l);
        if (rc) return rc;
    }
    return PFNAV_OK;
}
