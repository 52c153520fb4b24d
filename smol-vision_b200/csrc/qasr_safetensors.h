/*
 * qasr_safetensors.h - minimal read-only safetensors directory reader (host C).
 *
 * Replaces the role of the reference's mmap reader (qwen_asr_safetensors.c:194-228,
 * 309-371) for ONE purpose: hand tensor bytes to the one-time HBM upload
 * (north_star item 5).  mmap(PROT_READ) each *.safetensors shard in a directory,
 * parse the JSON header, look tensors up by name.
 */
#ifndef QASR_SAFETENSORS_H
#define QASR_SAFETENSORS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { QST_F32 = 0, QST_F16 = 1, QST_BF16 = 2, QST_OTHER = 3 };

typedef struct {
    char name[200];
    int dtype;
    int ndim;
    int64_t shape[8];
    const void *data; /* inside the mmap */
    size_t nbytes;
    size_t numel;
} qst_tensor_t;

typedef struct qst_dir qst_dir_t;

qst_dir_t *qst_open_dir(const char *model_dir);
/* view over caller-held tensors (entries copied, data not); close with qst_close */
qst_dir_t *qst_from_table(const qst_tensor_t *tensors, int n);
size_t qst_elem_size(int dtype);
void qst_close(qst_dir_t *d);
const qst_tensor_t *qst_find(const qst_dir_t *d, const char *name);
int qst_count(const qst_dir_t *d);
const qst_tensor_t *qst_at(const qst_dir_t *d, int i);

#ifdef __cplusplus
}
#endif
#endif
