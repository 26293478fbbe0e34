// C++17 host-side mirror of the reference crate's public surface, over the C ABI of fhe_precompiles_b200.h.
//
// The reference is a Rust crate; this image has no Rust toolchain, so the typed layer a Rust caller would use is provided in
// C++ (header-only, no dependencies beyond the C header and the shared library) with the reference's names, argument meaning
// and error behaviour, so that tests written against it read like the reference's own (tests/fheapp_test.cpp):
//
//   reference (Rust)                                   here (namespace fhe_precompiles)
//   lib.rs:3-44   enum FheError, i32 codes, strings    enum class FheError, error_code_to_str()
//   lib.rs:52     type PrecompileResult                PrecompileResult = Result<Bytes>
//   pack.rs:13-19 trait FHESerialize                   fhe_serialize() / T::fhe_deserialize() on every argument type
//   pack.rs:47-117 scalar / Vec<u8> impls              Unsigned64, Unsigned256, Signed, Fractional64, Bytes
//   pack.rs:21-45 Ciphertext / PublicKey impls         Ciphertext, PublicKey (opaque bincode holders: the bytes are parsed and
//                                                      validated inside the precompile, where the reference deserialises them)
//   pack.rs:119-266 pack_* / unpack_*                  pack::pack_one_argument ... pack::unpack_binary_operation
//   fhe.rs:161-779 FheApp::<49 precompiles>            FheApp::<same names>(input) -> PrecompileResult
//   fhe.rs:594,669,688 encrypt::<P> / reencrypt / decrypt   FheApp::encrypt<P>() / reencrypt<P>() / decrypt<P>()
//   testnet.rs:25 testnet::one::FHE                    testnet::one::FHE
//
// Every method is one call through the C symbol the reference exports for it (c_fhe.rs:74-141); all arithmetic runs in the
// CUDA engine behind that symbol.  There is no CPU fallback: without a GPU the calls return SunscreenError (code 7).
#pragma once

#include <array>
#include <cstdint>
#include <cstring>
#include <initializer_list>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#include "fhe_precompiles_b200.h"

namespace fhe_precompiles {

using Bytes = std::vector<uint8_t>;

// lib.rs:3-27: the variant <-> i32 mapping of `impl From<FheError> for i32`
enum class FheError : int32_t {
    UnexpectedEOF = 1,
    PlatformArchitecture = 2,
    InvalidEncoding = 3,
    Overflow = 4,
    FailedDecryption = 5,
    FailedEncryption = 6,
    SunscreenError = 7,
};

// lib.rs:30-44 FheError::error_code_to_str
inline const char *error_code_to_str(int32_t error_code) {
    switch (error_code) {
        case 1: return "Unexpected end of file";
        case 2: return "Platform architecture invalid";
        case 3: return "Invalid encoding";
        case 4: return "Overflow in FHE program";
        case 5: return "Invalid decryption";
        case 6: return "Invalid encryption";
        case 7: return "Base sunscreen error";
        default: return "Unknown error";
    }
}

// Result<T, FheError> with the handful of methods the reference's call sites use
template <class T>
class Result {
  public:
    Result(T value) : ok_(true), value_(std::move(value)), err_(FheError::SunscreenError) {}  // NOLINT: Ok(v)
    Result(FheError e) : ok_(false), value_(), err_(e) {}                                      // NOLINT: Err(e)
    bool is_ok() const { return ok_; }
    bool is_err() const { return !ok_; }
    explicit operator bool() const { return ok_; }
    // like Result::unwrap: a caller that did not check gets an exception instead of a panic
    const T &unwrap() const & {
        if (!ok_) throw std::runtime_error(std::string("called unwrap() on an Err value: ") + error_code_to_str((int32_t)err_));
        return value_;
    }
    T unwrap() && {
        if (!ok_) throw std::runtime_error(std::string("called unwrap() on an Err value: ") + error_code_to_str((int32_t)err_));
        return std::move(value_);
    }
    FheError unwrap_err() const {
        if (ok_) throw std::runtime_error("called unwrap_err() on an Ok value");
        return err_;
    }

  private:
    bool ok_;
    T value_;
    FheError err_;
};

using PrecompileResult = Result<Bytes>;  // lib.rs:52

// ---------------------------------------------------------------- argument types and their FHESerialize impls
namespace detail {
template <size_t N>
inline bool exact(const Bytes &b) {
    return b.size() == N;
}
inline uint64_t load_be64(const uint8_t *p) {
    uint64_t v = 0;
    for (int i = 0; i < 8; i++) v = (v << 8) | p[i];
    return v;
}
inline void store_be64(uint8_t *p, uint64_t v) {
    for (int i = 7; i >= 0; i--) {
        p[i] = (uint8_t)v;
        v >>= 8;
    }
}
}  // namespace detail

// pack.rs:47-59: 8 big-endian bytes; any other length is InvalidEncoding
struct Unsigned64 {
    uint64_t value = 0;
    Unsigned64() = default;
    Unsigned64(uint64_t v) : value(v) {}  // NOLINT: Unsigned64::from
    static Unsigned64 from(uint64_t v) { return Unsigned64(v); }
    bool operator==(const Unsigned64 &o) const { return value == o.value; }
    Bytes fhe_serialize() const {
        Bytes b(8);
        detail::store_be64(b.data(), value);
        return b;
    }
    static Result<Unsigned64> fhe_deserialize(const Bytes &b) {
        if (!detail::exact<8>(b)) return FheError::InvalidEncoding;
        return Unsigned64(detail::load_be64(b.data()));
    }
    static constexpr const char *kSuffix = "u64";
};

// pack.rs:61-73: 32 big-endian bytes
struct Unsigned256 {
    std::array<uint8_t, 32> be{};  // most significant byte first
    Unsigned256() = default;
    Unsigned256(uint64_t v) { detail::store_be64(be.data() + 24, v); }  // NOLINT: Unsigned256::from(u64)
    static Unsigned256 from(uint64_t v) { return Unsigned256(v); }
    bool operator==(const Unsigned256 &o) const { return be == o.be; }
    Bytes fhe_serialize() const { return Bytes(be.begin(), be.end()); }
    static Result<Unsigned256> fhe_deserialize(const Bytes &b) {
        if (!detail::exact<32>(b)) return FheError::InvalidEncoding;
        Unsigned256 v;
        std::memcpy(v.be.data(), b.data(), 32);
        return v;
    }
    static constexpr const char *kSuffix = "u256";
};

// pack.rs:75-89: i64 as 8 big-endian two's-complement bytes
struct Signed {
    int64_t value = 0;
    Signed() = default;
    Signed(int64_t v) : value(v) {}  // NOLINT: Signed::from
    static Signed from(int64_t v) { return Signed(v); }
    bool operator==(const Signed &o) const { return value == o.value; }
    Bytes fhe_serialize() const {
        Bytes b(8);
        detail::store_be64(b.data(), (uint64_t)value);
        return b;
    }
    static Result<Signed> fhe_deserialize(const Bytes &b) {
        if (!detail::exact<8>(b)) return FheError::InvalidEncoding;
        return Signed((int64_t)detail::load_be64(b.data()));
    }
    static constexpr const char *kSuffix = "i64";
};

// pack.rs:91-104: Fractional<64> as the 8 big-endian bytes of its f64 value
struct Fractional64 {
    double value = 0.0;
    Fractional64() = default;
    Fractional64(double v) : value(v) {}  // NOLINT: Fractional::<64>::from
    static Fractional64 from(double v) { return Fractional64(v); }
    bool operator==(const Fractional64 &o) const { return value == o.value; }
    Bytes fhe_serialize() const {
        uint64_t bits;
        std::memcpy(&bits, &value, 8);
        Bytes b(8);
        detail::store_be64(b.data(), bits);
        return b;
    }
    static Result<Fractional64> fhe_deserialize(const Bytes &b) {
        if (!detail::exact<8>(b)) return FheError::InvalidEncoding;
        const uint64_t bits = detail::load_be64(b.data());
        double v;
        std::memcpy(&v, &bits, 8);
        return Fractional64(v);
    }
    static constexpr const char *kSuffix = "frac64";
};

// pack.rs:106-117: Vec<u8> serialises as itself (the `public_data` argument of encrypt / reencrypt)
struct PublicData {
    Bytes bytes;
    PublicData() = default;
    PublicData(Bytes b) : bytes(std::move(b)) {}  // NOLINT
    PublicData(std::initializer_list<uint8_t> l) : bytes(l) {}
    Bytes fhe_serialize() const { return bytes; }
    static Result<PublicData> fhe_deserialize(const Bytes &b) { return PublicData(b); }
};

// pack.rs:21-45: bincode of sunscreen's Ciphertext / PublicKey.  Held as the serialised bytes: the precompile that receives
// them parses and validates them (InvalidEncoding / SunscreenError come back from there, as in the reference's fhe_binary_op).
struct Ciphertext {
    Bytes bincode;
    Ciphertext() = default;
    Ciphertext(Bytes b) : bincode(std::move(b)) {}  // NOLINT
    Bytes fhe_serialize() const { return bincode; }
    static Result<Ciphertext> fhe_deserialize(const Bytes &b) {
        if (b.empty()) return FheError::InvalidEncoding;
        return Ciphertext(b);
    }
};
struct PublicKey {
    Bytes bincode;
    PublicKey() = default;
    PublicKey(Bytes b) : bincode(std::move(b)) {}  // NOLINT
    Bytes fhe_serialize() const { return bincode; }
    static Result<PublicKey> fhe_deserialize(const Bytes &b) {
        if (b.empty()) return FheError::InvalidEncoding;
        return PublicKey(b);
    }
};

// ---------------------------------------------------------------- pack.rs:119-266
namespace pack {
constexpr size_t kIndexSize = 4;  // pack.rs:11 `type Index = u32`, big-endian on the wire

namespace detail {
inline void put_index(Bytes &out, size_t v) {
    out.push_back((uint8_t)(v >> 24));
    out.push_back((uint8_t)(v >> 16));
    out.push_back((uint8_t)(v >> 8));
    out.push_back((uint8_t)v);
}
inline size_t get_index(const Bytes &in, size_t at) {
    return ((size_t)in[at] << 24) | ((size_t)in[at + 1] << 16) | ((size_t)in[at + 2] << 8) | (size_t)in[at + 3];
}
inline Bytes slice(const Bytes &in, size_t from, size_t to) { return Bytes(in.begin() + (ptrdiff_t)from, in.begin() + (ptrdiff_t)to); }
}  // namespace detail

template <class A>
Bytes pack_one_argument(const A &a) {  // pack.rs:119-124
    return a.fhe_serialize();
}
template <class A>
Result<A> unpack_one_argument(const Bytes &input) {  // pack.rs:126-131
    return A::fhe_deserialize(input);
}
template <class A, class B>
Bytes pack_two_arguments(const A &a, const B &b) {  // pack.rs:133-151: [ix_1][a][b], ix_1 = 4 + len(a)
    const Bytes sa = a.fhe_serialize(), sb = b.fhe_serialize();
    Bytes out;
    out.reserve(kIndexSize + sa.size() + sb.size());
    detail::put_index(out, kIndexSize + sa.size());
    out.insert(out.end(), sa.begin(), sa.end());
    out.insert(out.end(), sb.begin(), sb.end());
    return out;
}
template <class A, class B>
Result<std::pair<A, B>> unpack_two_arguments(const Bytes &input) {  // pack.rs:153-175
    if (input.size() < kIndexSize) return FheError::UnexpectedEOF;
    const size_t ix1 = detail::get_index(input, 0);
    if (ix1 < kIndexSize || ix1 > input.size()) return FheError::UnexpectedEOF;  // the reference panics on such an offset
    auto a = A::fhe_deserialize(detail::slice(input, kIndexSize, ix1));
    if (!a) return a.unwrap_err();
    auto b = B::fhe_deserialize(detail::slice(input, ix1, input.size()));
    if (!b) return b.unwrap_err();
    return std::make_pair(std::move(a).unwrap(), std::move(b).unwrap());
}
inline Bytes pack_nullary_operation(const PublicKey &public_key) { return public_key.fhe_serialize(); }  // pack.rs:185-187
inline Result<PublicKey> unpack_nullary_operation(const Bytes &input) { return PublicKey::fhe_deserialize(input); }  // 197-199
template <class A, class B>
Bytes pack_binary_operation(const PublicKey &public_key, const A &a, const B &b) {  // pack.rs:208-231
    const Bytes pk = public_key.fhe_serialize(), sa = a.fhe_serialize(), sb = b.fhe_serialize();
    const size_t ix1 = pk.size() + 2 * kIndexSize, ix2 = ix1 + sa.size();
    Bytes out;
    out.reserve(ix2 + sb.size());
    detail::put_index(out, ix1);
    detail::put_index(out, ix2);
    out.insert(out.end(), pk.begin(), pk.end());
    out.insert(out.end(), sa.begin(), sa.end());
    out.insert(out.end(), sb.begin(), sb.end());
    return out;
}
template <class A, class B>
Result<std::tuple<PublicKey, A, B>> unpack_binary_operation(const Bytes &input) {  // pack.rs:238-266
    if (input.size() < 2 * kIndexSize) return FheError::UnexpectedEOF;
    const size_t ix1 = detail::get_index(input, 0), ix2 = detail::get_index(input, kIndexSize);
    if (ix1 < 2 * kIndexSize || ix2 < ix1 || ix2 > input.size()) return FheError::UnexpectedEOF;
    auto pk = PublicKey::fhe_deserialize(detail::slice(input, 2 * kIndexSize, ix1));
    if (!pk) return pk.unwrap_err();
    auto a = A::fhe_deserialize(detail::slice(input, ix1, ix2));
    if (!a) return a.unwrap_err();
    auto b = B::fhe_deserialize(detail::slice(input, ix2, input.size()));
    if (!b) return b.unwrap_err();
    return std::make_tuple(std::move(pk).unwrap(), std::move(a).unwrap(), std::move(b).unwrap());
}
}  // namespace pack

// ---------------------------------------------------------------- FheApp (fhe.rs:56-780)
class FheApp {
    using Symbol = int32_t (*)(const uint8_t *, size_t, uint8_t **, int64_t *);
    static PrecompileResult call(Symbol f, const Bytes &input) {
        uint8_t *out = nullptr;
        int64_t len = 0;
        const int32_t rc = f(input.empty() ? nullptr : input.data(), input.size(), &out, &len);
        if (rc != 0) return static_cast<FheError>(rc);
        Bytes r(out, out + len);
        fhe_free(out);
        return r;
    }

  public:
#define FHE_B200_METHOD(name) \
    PrecompileResult name(const Bytes &input) const { return call(&c_fhe_##name, input); }
#define FHE_B200_TYPE_METHODS(t)                 \
    FHE_B200_METHOD(add_cipher##t##_cipher##t)   \
    FHE_B200_METHOD(add_cipher##t##_##t)         \
    FHE_B200_METHOD(add_##t##_cipher##t)         \
    FHE_B200_METHOD(sub_cipher##t##_cipher##t)   \
    FHE_B200_METHOD(sub_cipher##t##_##t)         \
    FHE_B200_METHOD(sub_##t##_cipher##t)         \
    FHE_B200_METHOD(mul_cipher##t##_cipher##t)   \
    FHE_B200_METHOD(mul_cipher##t##_##t)         \
    FHE_B200_METHOD(mul_##t##_cipher##t)         \
    FHE_B200_METHOD(encrypt_##t)                 \
    FHE_B200_METHOD(reencrypt_##t)               \
    FHE_B200_METHOD(decrypt_##t)
    FHE_B200_TYPE_METHODS(u256)    // fhe.rs:161-268, 711, 735, 759
    FHE_B200_TYPE_METHODS(u64)     // fhe.rs:273-368, 717, 741, 765
    FHE_B200_TYPE_METHODS(i64)     // fhe.rs:373-468, 723, 747, 771
    FHE_B200_TYPE_METHODS(frac64)  // fhe.rs:469-576, 729, 753, 777
    FHE_B200_METHOD(public_key_bytes)  // fhe.rs:701-703 (the input is ignored)
#undef FHE_B200_TYPE_METHODS
#undef FHE_B200_METHOD

    // the generic forms of the threshold-network API (fhe.rs:594, 669, 688), dispatched on the plaintext type
    template <class P>
    PrecompileResult encrypt(const Bytes &input) const {
        return by_type<P>(&c_fhe_encrypt_u256, &c_fhe_encrypt_u64, &c_fhe_encrypt_i64, &c_fhe_encrypt_frac64, input);
    }
    template <class P>
    PrecompileResult reencrypt(const Bytes &input) const {
        return by_type<P>(&c_fhe_reencrypt_u256, &c_fhe_reencrypt_u64, &c_fhe_reencrypt_i64, &c_fhe_reencrypt_frac64, input);
    }
    template <class P>
    PrecompileResult decrypt(const Bytes &input) const {
        return by_type<P>(&c_fhe_decrypt_u256, &c_fhe_decrypt_u64, &c_fhe_decrypt_i64, &c_fhe_decrypt_frac64, input);
    }
    // the network public key as a typed value (the reference's `FHE.public_key`)
    PublicKey public_key() const { return PublicKey(public_key_bytes(Bytes()).unwrap()); }

  private:
    template <class P>
    static PrecompileResult by_type(Symbol u256, Symbol u64, Symbol i64, Symbol frac64, const Bytes &input) {
        if constexpr (std::is_same_v<P, Unsigned256>) return call(u256, input);
        else if constexpr (std::is_same_v<P, Unsigned64>) return call(u64, input);
        else if constexpr (std::is_same_v<P, Signed>) return call(i64, input);
        else {
            static_assert(std::is_same_v<P, Fractional64>, "plaintext type must be one of the reference's four");
            return call(frac64, input);
        }
    }
};

namespace testnet {
namespace one {
inline const FheApp FHE{};  // testnet.rs:25: the process-wide instance with the network keys (they live inside the library)
}  // namespace one
}  // namespace testnet

}  // namespace fhe_precompiles
