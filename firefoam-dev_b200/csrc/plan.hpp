// plan.hpp -- host-side, once-per-mesh preparation of the device row structure.
//
// Replaces the lazily-built parts of OpenFOAM's lduAddressing (ownerStartAddr, losortAddr,
// losortStartAddr; OF-dev lduAddressing.C) by ONE structure that serves
// Amul, sumA, negSumDiag, the DIC-class sweeps and the halo fix-up:
//
//   * an internal row order: natural (diagonal / none) or colour-major (DIC-class: greedy
//     multicolouring; DIC-exact: dependency levels of OpenFOAM's own face-order recurrences),
//   * a full-row sliced-ELL layout (slice height 32 = one warp; entry j of row r lives at
//     sliceBase[r/32] + 32*j + r%32, so a warp's j-th loads are one contiguous 128/256-byte
//     line), entries grouped [neighbours earlier in elimination order | later ones], each
//     group in ascending natural face order so that row sums are formed in exactly the order
//     of OpenFOAM's face loops,
//   * for every entry the natural face whose `upper` value it carries (value fill is a pure
//     gather, once per solve),
//   * the processor-interface rows as a CSR (row -> (patch-face slot) list) so that the halo
//     fix-up is a sorted-segment reduction, never an atomic.
//
// Pure C++ (no CUDA) so that tests can validate it on a CPU-only box.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace b200 {

enum class Ordering : int { Natural = 0, MultiColour = 1, Levels = 2 };

struct IfaceIn {
    int32_t nbrRank;
    int32_t nFaces;
    const int32_t* faceCells;
};

// Symmetric "single-read" Amul layout.  The matrix is stored ONCE: the later-in-order ("upper")
// entries of every row in a sliced ELL (value + column, coalesced).  A row's earlier ("lower")
// entries are 32-bit references (owner row a << 5 | q): the value is re-read from a's q-th upper
// entry, which was streamed moments earlier by a's own thread and is served by the 126 MB L2 on
// any bandwidth-reducing cell order (for the 16 M-cell hex box the reuse distance is < 5 MB).
// DRAM bytes per Amul: 16 F + 24 N + 4 N (row lengths) -- the algorithmic LDU minimum + 4 N --
// with no shared memory, no barriers, no atomics, and the row-sum order of OpenFOAM's face loop.
// Slices have a uniform width (global max row length) so that the position of an owner's entry
// is pure arithmetic (no offset-table lookup in the dependent-load chain); padded entries are
// never loaded, so the padding costs address space, not bandwidth.
struct SymPlan {
    bool valid = false;
    int32_t WU = 0, WL = 0;       // uniform slice widths: entry j of row r at (r/32)*32*W + 32*j + r%32
    int64_t nU = 0, nL = 0;       // padded entry counts (< 2^31: 32-bit positions on the device)
    std::vector<uint32_t> rowLen; // [N] nLower | nTotal<<16 with lower == smaller ROW INDEX (the full-row
                                  // ELL of a tiled multicolour plan splits by colour instead)
    std::vector<int32_t> uCol;    // [nU] column (padding: the row itself)
    std::vector<int32_t> uFace;   // [nU] natural face, -1 padding
    std::vector<uint32_t> lRef;   // [nL] (owner row << 5) | q
};

// Single-read layout for RENUMBERED natural plans ("SR": face-ordered).  A renumbered row must still add its
// faces in ascending NATURAL face order (bit-identical row sums), and in that order the entries whose value the
// row owns (neighbour has the larger row index) interleave with the entries it only references.  So the row keeps
// ONE list, aligned with the full-row ELL (same sliceBase / rowLen, ascending natural face order), of 32-bit words
//     meta = column << 5 | q        q == 31: own entry, the value is the row's next own value
//                                    q  < 31: reference to the q-th own value of row `column`
// and every coefficient is stored once, in a second sliced ELL of own values (per-slice width):
//     own value j of row r at  ownBase[r / 32] + 32 j + r % 32.
// DRAM bytes per Amul: 8 F (values) + 8 F (meta) + 28 N instead of the full-row ELL's 24 F + 28 N; the referenced
// values are L2 hits on a bandwidth-reduced order.  Replaces the "ranked" form of SymPlan, which staged the
// products by rank in shared memory and measured slower than the full-row ELL (profiles/r01_v7_dic_tiles.md).
struct SrPlan {
    bool valid = false;
    int64_t nOwn = 0;                 // padded own-value slots
    std::vector<uint32_t> meta;       // [nEntries] (padding: 0)
    std::vector<int64_t> ownBase;     // [nSlices + 1]
    std::vector<int32_t> ownFace;     // [nOwn] natural face of the slot, -1 padding
};

// Base cell order the row orders are derived from.  OpenFOAM meshes are normally bandwidth-reduced
// (renumberMesh), but nothing guarantees it: on a cache-hostile numbering a warp's neighbour gathers
// touch 32 different sectors per request and Amul drops to a quarter of the HBM roofline (measured on
// the block-shuffled polyhedral workload).  The plan may therefore renumber rows by reverse
// Cuthill-McKee.  Results do not change: every row still sums its faces in ascending natural face
// order (the order of OpenFOAM's face loop); only the order of the global dot-product sums differs.
enum class Renumber : int { Off = 0, Auto = -1, Force = 1 };

struct HostPlan {
    Ordering ordering = Ordering::Natural;
    bool renumbered = false;           // rows follow an RCM base order
    double spanNatural = 0, spanUsed = 0;   // mean |row(l) - row(u)| over faces, before / after
    double sectorsNatural = 0, sectorsUsed = 0;   // mean 32-B sectors per warp gather request (sampled)
    SymPlan sym;
    SrPlan sr;                         // renumbered natural plans only
    int32_t N = 0, F = 0;
    // row order
    std::vector<int32_t> perm;         // internal row -> natural cell  (empty == identity)
    std::vector<int32_t> iperm;        // natural cell -> internal row  (empty == identity)
    int32_t nColours = 1;
    std::vector<int32_t> colourStart;  // [nColours+1] cumulative rows per colour; the colour's row range
                                       // when nTiles == 1
    // Tiled multicolour order (DIC-class on large meshes): rows are ordered (tile of tileRows
    // consecutive base-order cells, colour, base position) instead of colour-major, so that a row and
    // its neighbours of the other colours stay a few thousand rows apart.  That keeps the symmetric
    // single-read Amul layout usable in the DIC-class modes (a row's earlier neighbours were streamed
    // moments ago: L2 hits), which a colour-major order (neighbours hundreds of MB upstream) does not.
    // The preconditioner is unchanged: elimination order = colour order, whatever the storage order.
    int32_t tileRows = 0;              // 0: one tile (plain colour-major)
    int32_t nTiles = 1;
    std::vector<int32_t> segStart;     // [nTiles*nColours + 1]: rows of (tile t, colour c) =
                                       // [segStart[t*nColours + c], segStart[t*nColours + c + 1])
    std::vector<int32_t> rowColour;    // [N] colour of every internal row (empty for Natural)
    // sliced ELL, both triangles
    int32_t nSlices = 0;
    int64_t nEntries = 0;              // padded
    std::vector<int64_t> sliceBase;    // [nSlices+1]
    std::vector<uint32_t> rowLen;      // [N]  nLower | nTotal<<16
    std::vector<int32_t> col;          // [nEntries] internal column (padding: the row itself)
    std::vector<int32_t> faceOf;       // [nEntries] natural face index, -1 for padding
    // 16-bit columns: the j-th neighbours of the 32 rows of a slice usually lie within 65 536 rows of
    // each other (banded / RCM / colour-major orders), so the kernels that are bound by the bytes of
    // the full-row ELL (Amul on permuted orders, DIC-class sweeps) read a 2-byte offset from a
    // per-(slice, j) base instead of the 4-byte column: 10 instead of 12 bytes per entry.
    // colBase[sliceBase[s]/32 + j] = smallest column of that slice entry, or -1 when the range does not
    // fit (the kernel then reads `col`).  Empty when no slice entry fits.
    std::vector<int32_t> colBase;      // [nEntries/32]
    std::vector<uint16_t> col16;       // [nEntries]
    double col16Fraction = 0;          // share of the slice entries that use the 16-bit form
    // interfaces (internal numbering); slot = position in the concatenation of all patches
    int32_t nIfaces = 0;
    std::vector<int32_t> nbrRank;      // [nIfaces]
    std::vector<int32_t> patchStart;   // [nIfaces+1] slot offsets
    std::vector<int32_t> slotRow;      // [nSlots] internal row of patch face (pack list)
    int32_t nBRows = 0;
    std::vector<int32_t> bRow;         // [nBRows] distinct interface rows, ascending
    std::vector<int32_t> bStart;       // [nBRows+1]
    std::vector<int32_t> bSlot;        // [nSlots] slots of each row, (patch, face) ascending
};

// Validates the LDU addressing (sizes, l<u, upper-triangular order) and builds the plan.
// Returns empty string on success, else an error message.
// sortColumns (MultiColour only): order the entries of each [earlier | later] group by COLUMN instead of by
// natural face index.  A DIC-class sweep has no OpenFOAM summation order to reproduce, and on a renumbered
// (RCM) mesh the face order of a row is unrelated to the positions of its neighbours: the j-th gathers of
// the 32 rows of a warp then scatter over 21-23 sectors per request; sorted by column they fall next to each
// other (11 sectors on the polyhedral workload -- tools/experiments/gather_sectors.py).
std::string build_plan(Ordering ordering, int32_t N, int32_t F, const int32_t* l, const int32_t* u,
                       int32_t nIfaces, const IfaceIn* ifaces, HostPlan& out,
                       Renumber renumber = Renumber::Off, int32_t tileRows = 0, bool sortColumns = false);

}  // namespace b200
