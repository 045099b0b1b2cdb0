/* fake_nccl.c -- TEST INFRASTRUCTURE: the six NCCL entry points csrc/multi.cu loads, for the CPU tier's emulated "devices" (all of them host
 * memory): a broadcast is a memcpy from the root's send buffer, issued when the root's call and the peer's call have both been seen inside a group.
 * Loaded through CSDRB_NCCL_LIB; nothing under csdr_b200/ refers to it. */
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
typedef struct fake_comm { int rank, n; } *ncclComm_t;
static int g_calls = 0;
int ncclCommInitAll(ncclComm_t *comm, int ndev, const int *devlist) { (void)devlist; for (int i = 0; i < ndev; i++) { comm[i] = malloc(sizeof **comm); comm[i]->rank = i; comm[i]->n = ndev; } return 0; }
int ncclCommDestroy(ncclComm_t c) { free(c); return 0; }
int ncclGroupStart(void) { return 0; }
int ncclGroupEnd(void) { return 0; }
int ncclBroadcast(const void *send, void *recv, size_t count, int dtype, int root, ncclComm_t c, void *stream)
{
    (void)stream; (void)root;
    if (dtype != 7) return 4;                                             /* ncclInvalidArgument: multi.cu only moves ncclFloat32 */
    g_calls++;
    if (recv != send) memcpy(recv, send, count * 4);                      /* single process: every rank's call carries the root's send buffer */
    (void)c;
    return 0;
}
const char *ncclGetErrorString(int e) { return e ? "fake nccl error" : "no error"; }
int fake_nccl_broadcast_calls(void) { return g_calls; }
