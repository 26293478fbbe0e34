/* A plain-C caller of the drop-in surface: what the reference's consumer (an EVM node binding c_fhe.rs through cgo or a
 * C shim) does, unchanged, against libfhe_precompiles_b200.so.  Reference interface: /root/reference/src/c_fhe.rs:23-71.
 *
 *   gcc -std=c11 -Wall -Iinclude examples/c_caller.c -Lfhe_precompiles_b200 -lfhe_precompiles_b200 \
 *       -Wl,-rpath,$PWD/fhe_precompiles_b200 -o c_caller
 *   ./c_caller packed_input.bin        (bytes produced by pack_binary_operation: public key, ciphertext a, ciphertext b)
 *
 * Without arguments it only asks for the network public key, which needs no GPU. */
#include <stdio.h>
#include <stdlib.h>

#include "fhe_precompiles_b200.h"

static unsigned char *read_file(const char *path, size_t *n) {
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long len = ftell(f);
    fseek(f, 0, SEEK_SET);
    unsigned char *buf = (unsigned char *)malloc(len > 0 ? (size_t)len : 1);
    *n = fread(buf, 1, (size_t)len, f);
    fclose(f);
    return buf;
}

int main(int argc, char **argv) {
    uint8_t *out = NULL;
    int64_t out_len = 0;
    int32_t rc = c_fhe_public_key_bytes(NULL, 0, &out, &out_len);
    if (rc != 0) {
        fprintf(stderr, "public_key_bytes: %s\n", fhe_error(rc));
        return 1;
    }
    printf("network public key: %lld bytes\n", (long long)out_len);
    fhe_free(out);
    if (argc < 2) return 0;

    size_t n = 0;
    unsigned char *packed = read_file(argv[1], &n);
    if (!packed) {
        fprintf(stderr, "cannot read %s\n", argv[1]);
        return 1;
    }
    rc = c_fhe_mul_cipheri64_cipheri64(packed, n, &out, &out_len);
    if (rc != 0) {
        /* code 7 carries a message of this library (e.g. "no CUDA device"); fhe_error is the reference's string */
        fprintf(stderr, "mul_cipheri64_cipheri64 failed: %d (%s) %s\n", rc, fhe_error(rc), fhe_b200_last_error());
        free(packed);
        return 2;
    }
    printf("result ciphertext: %lld bytes\n", (long long)out_len);
    fhe_free(out);

    /* the same call twice as one batch (extension): independent calls are sharded over the visible GPUs */
    fhe_b200_call calls[2];
    for (int i = 0; i < 2; i++) {
        calls[i].op = fhe_b200_op_index("mul_cipheri64_cipheri64");
        calls[i].bytes = packed;
        calls[i].bytes_length = n;
    }
    int64_t failed = fhe_b200_batch(calls, 2, 0);
    printf("batch: %lld failed, outputs %lld and %lld bytes\n", (long long)failed, (long long)calls[0].output_length,
           (long long)calls[1].output_length);
    for (int i = 0; i < 2; i++)
        if (calls[i].status == 0) fhe_free(calls[i].output);
    free(packed);
    return 0;
}
