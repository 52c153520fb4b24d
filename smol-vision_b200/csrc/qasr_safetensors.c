/* qasr_safetensors.c - see qasr_safetensors.h */
#include "qasr_safetensors.h"

#include <dirent.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#define QST_MAX_SHARDS 64

struct qst_dir {
    void *maps[QST_MAX_SHARDS];
    size_t map_len[QST_MAX_SHARDS];
    int n_maps;
    qst_tensor_t *tensors;
    int n, cap;
};

size_t qst_elem_size(int dtype) { return dtype == QST_F32 ? 4 : (dtype == QST_F16 || dtype == QST_BF16) ? 2 : 0; }

static void skip_ws(const char **p, const char *end) {
    while (*p < end && (**p == ' ' || **p == '\n' || **p == '\r' || **p == '\t')) (*p)++;
}

/* Parses a JSON string at *p (must point at the opening quote). */
static int parse_str(const char **p, const char *end, char *out, size_t cap) {
    if (*p >= end || **p != '"') return -1;
    (*p)++;
    size_t i = 0;
    while (*p < end && **p != '"') {
        char c = **p;
        if (c == '\\' && *p + 1 < end) { (*p)++; c = **p; }
        if (i + 1 < cap) out[i++] = c;
        (*p)++;
    }
    out[i] = 0;
    if (*p >= end) return -1;
    (*p)++;
    return 0;
}

/* Skips any JSON value (used for __metadata__ and unknown keys). */
static void skip_value(const char **p, const char *end) {
    skip_ws(p, end);
    if (*p >= end) return;
    if (**p == '"') { char tmp[8]; parse_str(p, end, tmp, sizeof tmp); return; }
    if (**p == '{' || **p == '[') {
        int depth = 0;
        while (*p < end) {
            char c = **p;
            if (c == '"') { char tmp[8]; parse_str(p, end, tmp, sizeof tmp); continue; }
            if (c == '{' || c == '[') depth++;
            if (c == '}' || c == ']') { depth--; if (depth == 0) { (*p)++; return; } }
            (*p)++;
        }
        return;
    }
    while (*p < end && **p != ',' && **p != '}' && **p != ']') (*p)++;
}

static int64_t parse_i64(const char **p, const char *end) {
    skip_ws(p, end);
    int64_t v = 0;
    while (*p < end && **p >= '0' && **p <= '9') { v = v * 10 + (**p - '0'); (*p)++; }
    return v;
}

static int parse_shard(qst_dir_t *d, const char *path) {
    int fd = open(path, O_RDONLY);
    if (fd < 0) return -1;
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 8) { close(fd); return -1; }
    void *m = mmap(NULL, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return -1;
    if (d->n_maps >= QST_MAX_SHARDS) { munmap(m, (size_t)st.st_size); return -1; }
    d->maps[d->n_maps] = m;
    d->map_len[d->n_maps] = (size_t)st.st_size;
    d->n_maps++;

    uint64_t hlen;
    memcpy(&hlen, m, 8);
    if (hlen > (uint64_t)st.st_size - 8) return -1; /* written so a huge hlen cannot wrap */
    const char *p = (const char *)m + 8, *end = p + hlen;
    const unsigned char *data0 = (const unsigned char *)m + 8 + hlen;

    skip_ws(&p, end);
    if (p >= end || *p != '{') return -1;
    p++;
    for (;;) {
        skip_ws(&p, end);
        if (p >= end || *p == '}') break;
        if (*p == ',') { p++; continue; }
        char name[200];
        if (parse_str(&p, end, name, sizeof name) != 0) return -1;
        skip_ws(&p, end);
        if (p >= end || *p != ':') return -1;
        p++;
        skip_ws(&p, end);
        if (strcmp(name, "__metadata__") == 0 || p >= end || *p != '{') { skip_value(&p, end); continue; }
        p++;
        qst_tensor_t t;
        memset(&t, 0, sizeof t);
        snprintf(t.name, sizeof t.name, "%s", name);
        t.dtype = QST_OTHER;
        size_t o0 = 0, o1 = 0;
        for (;;) {
            skip_ws(&p, end);
            if (p >= end) return -1;
            if (*p == '}') { p++; break; }
            if (*p == ',') { p++; continue; }
            char key[32];
            if (parse_str(&p, end, key, sizeof key) != 0) return -1;
            skip_ws(&p, end);
            if (p >= end || *p != ':') return -1;
            p++;
            skip_ws(&p, end);
            if (strcmp(key, "dtype") == 0) {
                char dt[16];
                if (parse_str(&p, end, dt, sizeof dt) != 0) return -1;
                t.dtype = !strcmp(dt, "F32") ? QST_F32 : !strcmp(dt, "F16") ? QST_F16 : !strcmp(dt, "BF16") ? QST_BF16 : QST_OTHER;
            } else if (strcmp(key, "shape") == 0 && *p == '[') {
                p++;
                for (;;) {
                    skip_ws(&p, end);
                    if (p >= end) return -1;
                    if (*p == ']') { p++; break; }
                    if (*p == ',') { p++; continue; }
                    if (t.ndim < 8) t.shape[t.ndim++] = parse_i64(&p, end); else parse_i64(&p, end);
                }
            } else if (strcmp(key, "data_offsets") == 0 && *p == '[') {
                p++;
                o0 = (size_t)parse_i64(&p, end);
                skip_ws(&p, end);
                if (p < end && *p == ',') p++;
                o1 = (size_t)parse_i64(&p, end);
                skip_ws(&p, end);
                if (p < end && *p == ']') p++;
            } else {
                skip_value(&p, end);
            }
        }
        if (o1 < o0 || o1 > (uint64_t)st.st_size - 8 - hlen) return -1;
        t.data = data0 + o0;
        t.nbytes = o1 - o0;
        t.numel = 1;
        for (int i = 0; i < t.ndim; i++) t.numel *= (size_t)t.shape[i];
        /* a header whose byte range disagrees with shape x dtype would make the upload read past the mmap */
        if (t.dtype != QST_OTHER && t.nbytes != t.numel * qst_elem_size(t.dtype)) return -1;
        if (d->n == d->cap) {
            d->cap = d->cap ? d->cap * 2 : 1024;
            d->tensors = (qst_tensor_t *)realloc(d->tensors, (size_t)d->cap * sizeof(qst_tensor_t));
            if (!d->tensors) return -1;
        }
        d->tensors[d->n++] = t;
    }
    return 0;
}

static int cmp_name(const void *a, const void *b) {
    return strcmp(((const qst_tensor_t *)a)->name, ((const qst_tensor_t *)b)->name);
}

static int cmp_path(const void *a, const void *b) { return strcmp(*(char *const *)a, *(char *const *)b); }

/* A directory view over tensors the caller already holds in memory (the reference's own mmap): nothing is mapped or
 * copied here, the entries are sorted by name for qst_find. */
qst_dir_t *qst_from_table(const qst_tensor_t *tensors, int n) {
    if (!tensors || n <= 0) return NULL;
    qst_dir_t *d = (qst_dir_t *)calloc(1, sizeof(qst_dir_t));
    if (!d) return NULL;
    d->tensors = (qst_tensor_t *)malloc((size_t)n * sizeof(qst_tensor_t));
    if (!d->tensors) { free(d); return NULL; }
    memcpy(d->tensors, tensors, (size_t)n * sizeof(qst_tensor_t));
    d->n = d->cap = n;
    qsort(d->tensors, (size_t)d->n, sizeof(qst_tensor_t), cmp_name);
    return d;
}

qst_dir_t *qst_open_dir(const char *model_dir) {
    DIR *dir = opendir(model_dir);
    if (!dir) return NULL;
    char *paths[QST_MAX_SHARDS];
    int np = 0;
    struct dirent *e;
    while ((e = readdir(dir)) != NULL && np < QST_MAX_SHARDS) {
        size_t l = strlen(e->d_name);
        if (l > 12 && strcmp(e->d_name + l - 12, ".safetensors") == 0) {
            size_t cap = strlen(model_dir) + l + 2;
            paths[np] = (char *)malloc(cap);
            snprintf(paths[np], cap, "%s/%s", model_dir, e->d_name);
            np++;
        }
    }
    closedir(dir);
    if (np == 0) return NULL;
    qsort(paths, (size_t)np, sizeof(char *), cmp_path);

    qst_dir_t *d = (qst_dir_t *)calloc(1, sizeof(qst_dir_t));
    int ok = d != NULL;
    for (int i = 0; i < np; i++) {
        if (ok && parse_shard(d, paths[i]) != 0) {
            fprintf(stderr, "qasr: cannot parse %s\n", paths[i]);
            ok = 0;
        }
        free(paths[i]);
    }
    if (!ok) { qst_close(d); return NULL; }
    qsort(d->tensors, (size_t)d->n, sizeof(qst_tensor_t), cmp_name);
    return d;
}

void qst_close(qst_dir_t *d) {
    if (!d) return;
    for (int i = 0; i < d->n_maps; i++) munmap(d->maps[i], d->map_len[i]);
    free(d->tensors);
    free(d);
}

const qst_tensor_t *qst_find(const qst_dir_t *d, const char *name) {
    int lo = 0, hi = d->n - 1;
    while (lo <= hi) {
        int mid = (lo + hi) / 2;
        int c = strcmp(d->tensors[mid].name, name);
        if (c == 0) return &d->tensors[mid];
        if (c < 0) lo = mid + 1; else hi = mid - 1;
    }
    return NULL;
}

int qst_count(const qst_dir_t *d) { return d->n; }
const qst_tensor_t *qst_at(const qst_dir_t *d, int i) { return (i >= 0 && i < d->n) ? &d->tensors[i] : NULL; }
