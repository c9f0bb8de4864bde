// avs_p2p.cuh -- layout of the peer-memory region shared by avs_dist.cu (setup, per-launch exchange kernels) and
// avs_cg.cu (the persistent CG kernel, which does its halo pushes and scalar all-reduces itself).
//
// Every rank exports one cudaMalloc region through CUDA IPC and maps the regions of all peers:
//   [ header: flags + mailboxes | p buffer 0 (owned rows, then halo slots) | p buffer 1 (same layout) ]
// The persistent CG kernel double-buffers p (p_{k+1} = z + beta p_k goes to the other buffer), so the rows a peer needs can
// be recomputed and stored into the peer's halo slots while the owner is still reading p_k; the per-launch path uses buffer 0.
#pragma once
#include <stdint.h>

#define P2P_MAX_RANKS 16
#define P2P_HEADER_BYTES 8192
struct P2PHeader {
    unsigned long long ready;                               // pull mode: sequence number of the p vector that is complete
    unsigned long long pad[15];
    unsigned long long flag[2][P2P_MAX_RANKS];              // mailbox sequence flags, written by the peers
    double mail[2][P2P_MAX_RANKS][4];                       // mailbox payload (<= 3 doubles used)
    unsigned long long pushFlag[P2P_MAX_RANKS];             // push mode: sender q's halo values of sequence n have landed
    // persistent CG kernel: flag-in-data mailboxes (the idea of NCCL's LL protocol) -- one 8-byte store carries 32 payload
    // bits and the 32-bit sequence number, so no fence separates data and flag: ll[parity][sender][value][half]
    unsigned long long ll[2][P2P_MAX_RANKS][3][2];
    unsigned long long localReady;                          // (unused since the arrival counter doubles as the local barrier)
    unsigned long long timedOut;                            // set by an exchange kernel that gave up waiting for a peer
};
static_assert(sizeof(P2PHeader) <= P2P_HEADER_BYTES, "header too large");

// What the persistent CG kernel needs to talk to its peers (passed by value).
struct PcgDist {
    int P = 1, myRank = 0;
    void *const *peerRegion = nullptr;   // device table of the mapped regions (peerRegion[myRank] = mine)
    const int32_t *sendIdx = nullptr;    // global row of every value a peer needs from me, grouped by peer
    const int2 *sendDst = nullptr;       // per send entry: (peer rank, element index in that peer's p vector)
    long long nSend = 0;
    long long rowBegin = 0;
    unsigned recvMask = 0, sendMask = 0; // peers I receive halo values from / push halo values to
    unsigned long long seqPush = 0, seqReduce = 0;  // last sequence numbers used before this launch
    unsigned long long pStrideBytes = 0;            // distance between p buffer 0 and p buffer 1 inside every rank's region
    const int32_t *boundarySlices = nullptr;        // slices of the local matrix that gather from a halo slot (ascending)
    long long nBoundary = 0;
    const uint8_t *sliceHalo = nullptr;             // per slice: 1 = in boundarySlices
};
