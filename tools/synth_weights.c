/*
 * synth_weights.c - deterministic synthetic Qwen3-ASR checkpoint writer.
 *
 * There are no real checkpoints offline, so parity and throughput are measured
 * on random-init weights of the named architecture (BASELINE.json north_star).
 * This tool writes `<dir>/model.safetensors` (all tensors BF16) plus a minimal
 * `vocab.json`, with the tensor names/shapes the reference loaders look up
 * (reference: qwen_asr_encoder.c:72-159, qwen_asr_decoder.c:55-158; variant
 * probe qwen_asr.c:146-203).  Values come from a counter-based hash so the
 * file is bit-identical on every machine (no libc/numpy RNG involved).
 *
 * Usage: synth_weights <0.6b|1.7b> <out_dir> [seed]
 */
#include <errno.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

typedef struct {
    char name[160];
    int ndim;
    int64_t shape[4];
    float mean, std;
    size_t numel, offset;
} tensor_spec_t;

static tensor_spec_t *g_specs = NULL;
static int g_n = 0, g_cap = 0;

static void add(const char *name, float mean, float std, int ndim, int64_t a, int64_t b, int64_t c, int64_t d) {
    if (g_n == g_cap) {
        g_cap = g_cap ? g_cap * 2 : 1024;
        g_specs = (tensor_spec_t *)realloc(g_specs, (size_t)g_cap * sizeof(tensor_spec_t));
    }
    tensor_spec_t *t = &g_specs[g_n++];
    memset(t, 0, sizeof(*t));
    snprintf(t->name, sizeof(t->name), "%s", name);
    t->ndim = ndim;
    t->shape[0] = a; t->shape[1] = b; t->shape[2] = c; t->shape[3] = d;
    t->mean = mean; t->std = std;
    t->numel = 1;
    for (int i = 0; i < ndim; i++) t->numel *= (size_t)t->shape[i];
}

static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

/* Irwin-Hall(4) approximation of N(0,1): bounded, cheap, reproducible. */
static inline float hash_normal(uint64_t key) {
    uint64_t h = splitmix64(key);
    float s = (float)(h & 0xFFFF) + (float)((h >> 16) & 0xFFFF) +
              (float)((h >> 32) & 0xFFFF) + (float)((h >> 48) & 0xFFFF);
    return (s * (1.0f / 65536.0f) - 2.0f) * 1.7320508f;
}

static inline uint16_t f32_to_bf16(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    u += 0x7FFFu + ((u >> 16) & 1u); /* round to nearest even */
    return (uint16_t)(u >> 16);
}

int main(int argc, char **argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s <0.6b|1.7b> <out_dir> [seed]\n", argv[0]);
        return 2;
    }
    int big = (strcmp(argv[1], "1.7b") == 0);
    if (!big && strcmp(argv[1], "0.6b") != 0) {
        fprintf(stderr, "unknown variant %s\n", argv[1]);
        return 2;
    }
    const char *dir = argv[2];
    uint64_t seed = argc > 3 ? strtoull(argv[3], NULL, 10) : 1234;

    int d = big ? 1024 : 896, F = big ? 4096 : 3584, enc_layers = big ? 24 : 18;
    int H = big ? 2048 : 1024, I = big ? 6144 : 3072, dec_layers = 28;
    int V = 151936, qd = 2048, kvd = 1024;
    char nm[192];
    const char *E = "thinker.audio_tower.";

    snprintf(nm, sizeof nm, "%sconv2d1.weight", E); add(nm, 0, 0.5f, 4, 480, 1, 3, 3);
    snprintf(nm, sizeof nm, "%sconv2d1.bias", E);   add(nm, 0, 0.05f, 1, 480, 0, 0, 0);
    for (int c = 2; c <= 3; c++) {
        snprintf(nm, sizeof nm, "%sconv2d%d.weight", E, c); add(nm, 0, 1.6f / sqrtf(4320.f), 4, 480, 480, 3, 3);
        snprintf(nm, sizeof nm, "%sconv2d%d.bias", E, c);   add(nm, 0, 0.05f, 1, 480, 0, 0, 0);
    }
    snprintf(nm, sizeof nm, "%sconv_out.weight", E); add(nm, 0, 1.6f / sqrtf(7680.f), 2, d, 7680, 0, 0);
    for (int l = 0; l < enc_layers; l++) {
        const char *pn[4] = {"q_proj", "k_proj", "v_proj", "out_proj"};
        for (int p = 0; p < 4; p++) {
            float g = (p == 3) ? 0.5f : 1.0f;
            snprintf(nm, sizeof nm, "%slayers.%d.self_attn.%s.weight", E, l, pn[p]); add(nm, 0, g / sqrtf((float)d), 2, d, d, 0, 0);
            snprintf(nm, sizeof nm, "%slayers.%d.self_attn.%s.bias", E, l, pn[p]);   add(nm, 0, 0.02f, 1, d, 0, 0, 0);
        }
        snprintf(nm, sizeof nm, "%slayers.%d.self_attn_layer_norm.weight", E, l); add(nm, 1, 0.05f, 1, d, 0, 0, 0);
        snprintf(nm, sizeof nm, "%slayers.%d.self_attn_layer_norm.bias", E, l);   add(nm, 0, 0.02f, 1, d, 0, 0, 0);
        snprintf(nm, sizeof nm, "%slayers.%d.fc1.weight", E, l); add(nm, 0, 1.0f / sqrtf((float)d), 2, F, d, 0, 0);
        snprintf(nm, sizeof nm, "%slayers.%d.fc1.bias", E, l);   add(nm, 0, 0.02f, 1, F, 0, 0, 0);
        snprintf(nm, sizeof nm, "%slayers.%d.fc2.weight", E, l); add(nm, 0, 0.5f / sqrtf((float)F), 2, d, F, 0, 0);
        snprintf(nm, sizeof nm, "%slayers.%d.fc2.bias", E, l);   add(nm, 0, 0.02f, 1, d, 0, 0, 0);
        snprintf(nm, sizeof nm, "%slayers.%d.final_layer_norm.weight", E, l); add(nm, 1, 0.05f, 1, d, 0, 0, 0);
        snprintf(nm, sizeof nm, "%slayers.%d.final_layer_norm.bias", E, l);   add(nm, 0, 0.02f, 1, d, 0, 0, 0);
    }
    snprintf(nm, sizeof nm, "%sln_post.weight", E); add(nm, 1, 0.05f, 1, d, 0, 0, 0);
    snprintf(nm, sizeof nm, "%sln_post.bias", E);   add(nm, 0, 0.02f, 1, d, 0, 0, 0);
    snprintf(nm, sizeof nm, "%sproj1.weight", E);   add(nm, 0, 1.0f / sqrtf((float)d), 2, d, d, 0, 0);
    snprintf(nm, sizeof nm, "%sproj1.bias", E);     add(nm, 0, 0.02f, 1, d, 0, 0, 0);
    snprintf(nm, sizeof nm, "%sproj2.weight", E);   add(nm, 0, 0.12f / sqrtf((float)d), 2, H, d, 0, 0);
    snprintf(nm, sizeof nm, "%sproj2.bias", E);     add(nm, 0, 0.005f, 1, H, 0, 0, 0);

    const char *D = "thinker.model.";
    snprintf(nm, sizeof nm, "%sembed_tokens.weight", D); add(nm, 0, 0.04f, 2, V, H, 0, 0);
    for (int l = 0; l < dec_layers; l++) {
        snprintf(nm, sizeof nm, "%slayers.%d.self_attn.q_proj.weight", D, l); add(nm, 0, 1.0f / sqrtf((float)H), 2, qd, H, 0, 0);
        snprintf(nm, sizeof nm, "%slayers.%d.self_attn.k_proj.weight", D, l); add(nm, 0, 1.0f / sqrtf((float)H), 2, kvd, H, 0, 0);
        snprintf(nm, sizeof nm, "%slayers.%d.self_attn.v_proj.weight", D, l); add(nm, 0, 1.0f / sqrtf((float)H), 2, kvd, H, 0, 0);
        snprintf(nm, sizeof nm, "%slayers.%d.self_attn.o_proj.weight", D, l); add(nm, 0, 0.3f / sqrtf((float)qd), 2, H, qd, 0, 0);
        snprintf(nm, sizeof nm, "%slayers.%d.self_attn.q_norm.weight", D, l); add(nm, 1, 0.05f, 1, 128, 0, 0, 0);
        snprintf(nm, sizeof nm, "%slayers.%d.self_attn.k_norm.weight", D, l); add(nm, 1, 0.05f, 1, 128, 0, 0, 0);
        snprintf(nm, sizeof nm, "%slayers.%d.input_layernorm.weight", D, l);  add(nm, 1, 0.05f, 1, H, 0, 0, 0);
        snprintf(nm, sizeof nm, "%slayers.%d.post_attention_layernorm.weight", D, l); add(nm, 1, 0.05f, 1, H, 0, 0, 0);
        snprintf(nm, sizeof nm, "%slayers.%d.mlp.gate_proj.weight", D, l); add(nm, 0, 1.0f / sqrtf((float)H), 2, I, H, 0, 0);
        snprintf(nm, sizeof nm, "%slayers.%d.mlp.up_proj.weight", D, l);   add(nm, 0, 1.0f / sqrtf((float)H), 2, I, H, 0, 0);
        snprintf(nm, sizeof nm, "%slayers.%d.mlp.down_proj.weight", D, l); add(nm, 0, 1.0f / sqrtf((float)I), 2, H, I, 0, 0);
    }
    snprintf(nm, sizeof nm, "%snorm.weight", D); add(nm, 1, 0.05f, 1, H, 0, 0, 0);

    /* Header JSON */
    size_t off = 0;
    for (int i = 0; i < g_n; i++) { g_specs[i].offset = off; off += g_specs[i].numel * 2; }
    size_t total_data = off;
    size_t hcap = (size_t)g_n * 320 + 64, hl = 0;
    char *hdr = (char *)malloc(hcap);
    hl += (size_t)snprintf(hdr + hl, hcap - hl, "{");
    for (int i = 0; i < g_n; i++) {
        tensor_spec_t *t = &g_specs[i];
        hl += (size_t)snprintf(hdr + hl, hcap - hl, "%s\"%s\":{\"dtype\":\"BF16\",\"shape\":[", i ? "," : "", t->name);
        for (int k = 0; k < t->ndim; k++)
            hl += (size_t)snprintf(hdr + hl, hcap - hl, "%s%lld", k ? "," : "", (long long)t->shape[k]);
        hl += (size_t)snprintf(hdr + hl, hcap - hl, "],\"data_offsets\":[%zu,%zu]}", t->offset, t->offset + t->numel * 2);
    }
    hl += (size_t)snprintf(hdr + hl, hcap - hl, "}");
    while (hl % 8) hdr[hl++] = ' ';

    mkdir(dir, 0755);
    char path[1024];
    snprintf(path, sizeof path, "%s/model.safetensors", dir);
    FILE *f = fopen(path, "wb");
    if (!f) { fprintf(stderr, "cannot open %s: %s\n", path, strerror(errno)); return 1; }
    uint64_t hl64 = hl;
    fwrite(&hl64, 8, 1, f);
    fwrite(hdr, 1, hl, f);

    size_t chunk = 1u << 22;
    uint16_t *buf = (uint16_t *)malloc(chunk * 2);
    for (int i = 0; i < g_n; i++) {
        tensor_spec_t *t = &g_specs[i];
        uint64_t base = splitmix64(seed * 0x100000001B3ULL + (uint64_t)i) << 1;
        for (size_t s = 0; s < t->numel; s += chunk) {
            size_t n = t->numel - s < chunk ? t->numel - s : chunk;
#pragma omp parallel for schedule(static)
            for (size_t j = 0; j < n; j++)
                buf[j] = f32_to_bf16(t->mean + t->std * hash_normal(base + s + j));
            if (fwrite(buf, 2, n, f) != n) { fprintf(stderr, "short write\n"); return 1; }
        }
    }
    fclose(f);

    snprintf(path, sizeof path, "%s/vocab.json", dir);
    f = fopen(path, "w");
    if (f) { fprintf(f, "{\"a\": 0, \"b\": 1, \"c\": 2}\n"); fclose(f); }
    fprintf(stderr, "synth_weights: %s seed=%llu tensors=%d bytes=%zu -> %s\n", argv[1],
            (unsigned long long)seed, g_n, total_data + 8 + hl, dir);
    return 0;
}
