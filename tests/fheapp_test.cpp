// The reference's own tests (/root/reference/src/fhe.rs:1038-2303, src/pack.rs tests) restated against the C++ mirror of its
// surface (include/fhe_precompiles_b200.hpp): same names, same values (16 and 4 -> 20 / 12 / 64), same known answers.
// Where the reference reaches into `FHE.runtime` (generate_keys, encrypt, decrypt with a generated key) this file uses the
// precompiles that do the same through the drop-in surface: encrypt_* under the network key, decrypt_* with the network key.
//
//   fheapp_test --cpu             host-only part (framing, scalar encodings, error strings, public_key_bytes): no GPU needed
//   fheapp_test <tests/data dir>  everything; needs a B200
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <string>

#include "fhe_precompiles_b200.hpp"

using namespace fhe_precompiles;
using fhe_precompiles::testnet::one::FHE;
using namespace fhe_precompiles::pack;

static int g_checks = 0;
#define CHECK(...)                                                                   \
    do {                                                                             \
        g_checks++;                                                                  \
        if (!(__VA_ARGS__)) {                                                               \
            std::fprintf(stderr, "%s:%d: check failed: %s (%s)\n", __FILE__, __LINE__, #__VA_ARGS__, fhe_b200_last_error()); \
            std::exit(1);                                                            \
        }                                                                            \
    } while (0)

static Bytes sha512(const Bytes &b) {
    Bytes out(64);
    fhe_b200_sha512(b.data(), b.size(), 0, out.data());
    return out;
}
static Bytes read_file(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    return Bytes(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
}

// ---------------------------------------------------------------- host-only
static void pack_round_trips() {  // pack.rs tests: what goes in comes out, offsets are big-endian u32
    const Unsigned64 a(16);
    const PublicData d{1, 2, 3};
    const Bytes two = pack_two_arguments(a, d);
    CHECK(two.size() == 4 + 8 + 3 && two[0] == 0 && two[1] == 0 && two[2] == 0 && two[3] == 12);
    auto back = unpack_two_arguments<Unsigned64, PublicData>(two);
    CHECK(back.is_ok() && back.unwrap().first == a && back.unwrap().second.bytes == d.bytes);
    const PublicKey pk(Bytes{9, 9, 9, 9, 9});
    const Bytes three = pack_binary_operation(pk, Signed(-5), Fractional64(0.25));
    CHECK(three.size() == 8 + 5 + 8 + 8 && three[3] == 13 && three[7] == 21);
    auto b3 = unpack_binary_operation<Signed, Fractional64>(three);
    CHECK(b3.is_ok() && std::get<0>(b3.unwrap()).bincode == pk.bincode && std::get<1>(b3.unwrap()) == Signed(-5) &&
          std::get<2>(b3.unwrap()) == Fractional64(0.25));
    CHECK(unpack_one_argument<Unsigned256>(pack_one_argument(Unsigned256(7))).unwrap() == Unsigned256(7));
    // too short for its offsets -> UnexpectedEOF (pack.rs:159-161, 244-246)
    CHECK(unpack_two_arguments<Unsigned64, PublicData>(Bytes{0, 0}).unwrap_err() == FheError::UnexpectedEOF);
    CHECK((unpack_binary_operation<Signed, Signed>(Bytes{0, 0, 0, 8, 0, 0}).unwrap_err() == FheError::UnexpectedEOF));
    // a scalar of the wrong length -> InvalidEncoding (pack.rs:56, 70, 86, 101)
    CHECK(Unsigned64::fhe_deserialize(Bytes(7)).unwrap_err() == FheError::InvalidEncoding);
    CHECK(Unsigned256::fhe_deserialize(Bytes(8)).unwrap_err() == FheError::InvalidEncoding);
    CHECK(Signed::fhe_deserialize(Bytes(9)).unwrap_err() == FheError::InvalidEncoding);
    CHECK(Fractional64::fhe_deserialize(Bytes()).unwrap_err() == FheError::InvalidEncoding);
    // two's complement, IEEE bits, most significant byte first
    CHECK((Signed(-2).fhe_serialize() == Bytes{0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xfe}));
    CHECK((Fractional64(1.0).fhe_serialize() == Bytes{0x3f, 0xf0, 0, 0, 0, 0, 0, 0}));
    CHECK(Unsigned256(0x0102).fhe_serialize()[30] == 1 && Unsigned256(0x0102).fhe_serialize()[31] == 2);
}
static void error_strings() {  // lib.rs:30-44, c_fhe.rs:66-71
    for (int32_t c = 0; c <= 8; c++) CHECK(std::string(error_code_to_str(c)) == fhe_error(c));
    CHECK((int32_t)FheError::FailedDecryption == 5 && (int32_t)FheError::FailedEncryption == 6);
}
static void public_key_bytes_works() {  // fhe.rs:701-703: any input, same key; the framing helpers are the identity on it
    const Bytes k = FHE.public_key_bytes(Bytes{1, 2, 3}).unwrap();
    CHECK(k.size() == 410994 && k == FHE.public_key_bytes(Bytes()).unwrap());
    CHECK(unpack_nullary_operation(pack_nullary_operation(PublicKey(k))).unwrap().bincode == k);
}
// ---------------------------------------------------------------- GPU (the engine refuses every call without a device: code 7)
static void framing_errors_reach_the_caller() {  // fhe_binary_op: unpack comes first (pack.rs:244-246)
    CHECK(FHE.add_cipheru64_cipheru64(Bytes{0, 0, 0}).unwrap_err() == FheError::UnexpectedEOF);
    CHECK(FHE.encrypt_u64(Bytes{0}).unwrap_err() == FheError::UnexpectedEOF);
    CHECK(FHE.mul_cipheri64_i64(Bytes(8, 0xff)).unwrap_err() == FheError::UnexpectedEOF);  // offsets outside the buffer
}

template <class P>
static Ciphertext enc(const P &v, uint8_t salt) {
    return Ciphertext(FHE.encrypt<P>(pack_two_arguments(v, PublicData{1, 2, 3, salt})).unwrap());
}
template <class P>
static P dec(const Bytes &ct) {
    return P::fhe_deserialize(FHE.decrypt<P>(ct).unwrap()).unwrap();
}
// precompile_fhe_op_works (fhe.rs:2309-2338)
template <class A, class B, class C, class F>
static void precompile_fhe_op_works(F fhe_op, const PublicKey &public_key, const A &a, const B &b, const C &expected) {
    const Bytes input = pack_binary_operation(public_key, a, b);
    const Bytes output = fhe_op(input).unwrap();
    CHECK(dec<C>(output) == expected);
}
#define OPS_FOR(t, T)                                                                                                          \
    static void precompiles_##t##_work(const PublicKey &pk) {                                                                  \
        const T a(16), b(4);                                                                                                   \
        const Ciphertext ea = enc(a, 1), eb = enc(b, 2);                                                                       \
        precompile_fhe_op_works([](const Bytes &i) { return FHE.add_cipher##t##_cipher##t(i); }, pk, ea, eb, T(20));           \
        precompile_fhe_op_works([](const Bytes &i) { return FHE.add_cipher##t##_##t(i); }, pk, ea, b, T(20));                  \
        precompile_fhe_op_works([](const Bytes &i) { return FHE.add_##t##_cipher##t(i); }, pk, a, eb, T(20));                  \
        precompile_fhe_op_works([](const Bytes &i) { return FHE.sub_cipher##t##_cipher##t(i); }, pk, ea, eb, T(12));           \
        precompile_fhe_op_works([](const Bytes &i) { return FHE.sub_cipher##t##_##t(i); }, pk, ea, b, T(12));                  \
        precompile_fhe_op_works([](const Bytes &i) { return FHE.sub_##t##_cipher##t(i); }, pk, a, eb, T(12));                  \
        precompile_fhe_op_works([](const Bytes &i) { return FHE.mul_cipher##t##_cipher##t(i); }, pk, ea, eb, T(64));           \
        precompile_fhe_op_works([](const Bytes &i) { return FHE.mul_cipher##t##_##t(i); }, pk, ea, b, T(64));                  \
        precompile_fhe_op_works([](const Bytes &i) { return FHE.mul_##t##_cipher##t(i); }, pk, a, eb, T(64));                  \
    }
OPS_FOR(u256, Unsigned256)
OPS_FOR(u64, Unsigned64)
OPS_FOR(i64, Signed)
OPS_FOR(frac64, Fractional64)

static void fhe_encrypt_test() {  // fhe.rs:2083-2121 (Linux known answer)
    const Unsigned256 value(12);
    const Bytes input = pack_two_arguments(value, PublicData{1, 2, 3});
    const Bytes result = FHE.encrypt<Unsigned256>(input).unwrap();
    CHECK(dec<Unsigned256>(result) == value);
    const Bytes want{190, 214, 153, 167, 205, 130, 61,  102, 188, 80,  220, 159, 38,  110, 126, 216, 148, 46, 220, 80,  18, 189,
                     177, 187, 108, 99,  32,  72,  250, 225, 2,   166, 33,  155, 22,  86,  221, 82,  4,   174, 144, 196, 45, 28,
                     190, 100, 194, 192, 37,  81,  203, 227, 46,  179, 59,  153, 20,  118, 191, 69,  244, 113, 180, 123};
    CHECK(sha512(result) == want);
}
static void encrypt_same_seed_and_value_works() {  // fhe.rs:2124-2140: a - a is a transparent ciphertext and must be accepted
    const Bytes input = pack_two_arguments(Unsigned256(16), PublicData{1, 2, 3, 4});
    const Ciphertext a = Ciphertext::fhe_deserialize(FHE.encrypt<Unsigned256>(input).unwrap()).unwrap();
    const Ciphertext b = Ciphertext::fhe_deserialize(FHE.encrypt<Unsigned256>(input).unwrap()).unwrap();
    CHECK(a.bincode == b.bincode);
    const Bytes result = FHE.sub_cipheru256_cipheru256(pack_binary_operation(FHE.public_key(), a, b)).unwrap();
    CHECK(Unsigned256::fhe_deserialize(FHE.decrypt_u256(result).unwrap()).unwrap() == Unsigned256(0));
}
static void fhe_decrypt_test() {  // fhe.rs:2248-2303
    const PublicData public_data{1, 2, 3};
    CHECK(dec<Unsigned256>(FHE.encrypt<Unsigned256>(pack_two_arguments(Unsigned256(12), public_data)).unwrap()) == Unsigned256(12));
    CHECK(dec<Unsigned64>(FHE.encrypt<Unsigned64>(pack_two_arguments(Unsigned64(12), public_data)).unwrap()) == Unsigned64(12));
    CHECK(dec<Signed>(FHE.encrypt<Signed>(pack_two_arguments(Signed(12), public_data)).unwrap()) == Signed(12));
    CHECK(dec<Fractional64>(FHE.encrypt<Fractional64>(pack_two_arguments(Fractional64(12.0), public_data)).unwrap()) ==
          Fractional64(12.0));
    CHECK(dec<Signed>(FHE.encrypt<Signed>(pack_two_arguments(Signed(-12), public_data)).unwrap()) == Signed(-12));
}
static void fhe_reencrypt_test(const std::string &data_dir) {  // fhe.rs:2188-2246 (Linux known answer)
    const PublicKey public_key(read_file(data_dir + "/public_key.bin"));
    CHECK(public_key.bincode.size() > 400000);
    const Unsigned256 value(12);
    const PublicData public_data{1, 2, 3};
    const Bytes result = FHE.encrypt<Unsigned256>(pack_two_arguments(value, public_data)).unwrap();
    const Ciphertext ciphertext(result);
    CHECK(dec<Unsigned256>(result) == value);
    const Bytes re = FHE.reencrypt<Unsigned256>(pack_binary_operation(public_key, ciphertext, public_data)).unwrap();
    const Bytes want{130, 189, 175, 155, 159, 130, 159, 220, 70,  102, 26,  228, 211, 59,  132, 240, 108, 2,   240, 176, 42, 236,
                     90,  30,  232, 41,  62,  25,  27,  239, 158, 39,  224, 40,  62,  212, 113, 151, 199, 5,   155, 15,  9,  35,
                     77,  46,  238, 46,  133, 185, 243, 242, 89,  101, 121, 56,  85,  103, 101, 0,   201, 200, 182, 64};
    CHECK(sha512(re) == want);
    // re-encrypting under the network key itself gives a ciphertext the network can decrypt again (fhe_refresh_test's flow)
    const Bytes again = FHE.reencrypt<Unsigned256>(pack_binary_operation(FHE.public_key(), ciphertext, public_data)).unwrap();
    CHECK(dec<Unsigned256>(again) == value);
}
static void errors_through_the_engine() {
    const PublicKey pk = FHE.public_key();
    const Ciphertext a = enc(Unsigned64(16), 7);
    // a scalar of the wrong width where the precompile expects its plaintext type -> InvalidEncoding (pack.rs:56)
    CHECK(FHE.add_cipheru64_u64(pack_binary_operation(pk, a, Unsigned256(4))).unwrap_err() == FheError::InvalidEncoding);
    // garbage where a ciphertext is expected -> InvalidEncoding (pack.rs:30)
    CHECK(FHE.add_cipheru64_cipheru64(pack_binary_operation(pk, a, PublicData{1, 2, 3})).unwrap_err() == FheError::InvalidEncoding);
    // a ciphertext of another plaintext type -> SunscreenError (fhe.rs:28: the runtime rejects the argument type)
    CHECK(FHE.add_cipheri64_cipheri64(pack_binary_operation(pk, a, a)).unwrap_err() == FheError::SunscreenError);
}

int main(int argc, char **argv) {
    pack_round_trips();
    error_strings();
    public_key_bytes_works();
    if (argc > 1 && std::string(argv[1]) == "--cpu") {
        std::printf("ok cpu %d checks\n", g_checks);
        return 0;
    }
    if (argc < 2) {
        std::fprintf(stderr, "usage: fheapp_test --cpu | <tests/data directory>\n");
        return 2;
    }
    framing_errors_reach_the_caller();
    const PublicKey pk = FHE.public_key();
    precompiles_u256_work(pk);
    precompiles_u64_work(pk);
    precompiles_i64_work(pk);
    precompiles_frac64_work(pk);
    fhe_encrypt_test();
    encrypt_same_seed_and_value_works();
    fhe_decrypt_test();
    fhe_reencrypt_test(argv[1]);
    errors_through_the_engine();
    std::printf("ok gpu %d checks\n", g_checks);
    return 0;
}
