//! FFI shim over `libfhe_precompiles_b200.so` with the reference's surface: `FheApp::<precompile>(&[u8]) -> PrecompileResult`
//! (reference: src/fhe.rs:161-779, error codes src/lib.rs:14-27). Not compiled in this repository's image.

#[derive(Debug, Clone, PartialEq, Eq)]
pub enum FheError {
    UnexpectedEOF,
    PlatformArchitecture,
    InvalidEncoding,
    Overflow,
    FailedDecryption,
    FailedEncryption,
    EngineError(String),
}

pub type PrecompileResult = Result<Vec<u8>, FheError>;

type Precompile = unsafe extern "C" fn(*const u8, libc::size_t, *mut *mut u8, *mut i64) -> i32;

#[link(name = "fhe_precompiles_b200")]
extern "C" {
    pub fn c_fhe_add_cipheru256_cipheru256(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_add_cipheru256_u256(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_add_u256_cipheru256(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_sub_cipheru256_cipheru256(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_sub_cipheru256_u256(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_sub_u256_cipheru256(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_mul_cipheru256_cipheru256(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_mul_cipheru256_u256(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_mul_u256_cipheru256(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_add_cipheru64_cipheru64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_add_cipheru64_u64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_add_u64_cipheru64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_sub_cipheru64_cipheru64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_sub_cipheru64_u64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_sub_u64_cipheru64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_mul_cipheru64_cipheru64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_mul_cipheru64_u64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_mul_u64_cipheru64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_add_cipheri64_cipheri64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_add_cipheri64_i64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_add_i64_cipheri64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_sub_cipheri64_cipheri64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_sub_cipheri64_i64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_sub_i64_cipheri64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_mul_cipheri64_cipheri64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_mul_cipheri64_i64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_mul_i64_cipheri64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_add_cipherfrac64_cipherfrac64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_add_cipherfrac64_frac64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_add_frac64_cipherfrac64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_sub_cipherfrac64_cipherfrac64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_sub_cipherfrac64_frac64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_sub_frac64_cipherfrac64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_mul_cipherfrac64_cipherfrac64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_mul_cipherfrac64_frac64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_mul_frac64_cipherfrac64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_encrypt_u256(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_encrypt_u64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_encrypt_i64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_encrypt_frac64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_reencrypt_u256(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_reencrypt_u64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_reencrypt_i64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_reencrypt_frac64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_decrypt_u256(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_decrypt_u64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_decrypt_i64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_decrypt_frac64(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn c_fhe_public_key_bytes(bytes: *const u8, len: libc::size_t, out: *mut *mut u8, out_len: *mut i64) -> i32;
    pub fn fhe_free(bytes: *const u8);
    pub fn fhe_error(code: i32) -> *const libc::c_char;
    pub fn fhe_b200_last_error() -> *const libc::c_char;
}

fn call(f: Precompile, input: &[u8]) -> PrecompileResult {
    let mut out: *mut u8 = std::ptr::null_mut();
    let mut len: i64 = 0;
    let rc = unsafe { f(input.as_ptr(), input.len(), &mut out, &mut len) };
    match rc {
        0 => {
            let v = unsafe { std::slice::from_raw_parts(out, len as usize) }.to_vec();
            unsafe { fhe_free(out) };
            Ok(v)
        }
        1 => Err(FheError::UnexpectedEOF),
        2 => Err(FheError::PlatformArchitecture),
        3 => Err(FheError::InvalidEncoding),
        4 => Err(FheError::Overflow),
        5 => Err(FheError::FailedDecryption),
        6 => Err(FheError::FailedEncryption),
        _ => Err(FheError::EngineError(
            unsafe { std::ffi::CStr::from_ptr(fhe_b200_last_error()) }.to_string_lossy().into_owned(),
        )),
    }
}

/// Same role as the reference's `FheApp`; stateless here because the engine is process-global inside the library.
pub struct FheApp;

impl FheApp {
    pub fn add_cipheru256_cipheru256(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_add_cipheru256_cipheru256, input) }
    pub fn add_cipheru256_u256(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_add_cipheru256_u256, input) }
    pub fn add_u256_cipheru256(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_add_u256_cipheru256, input) }
    pub fn sub_cipheru256_cipheru256(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_sub_cipheru256_cipheru256, input) }
    pub fn sub_cipheru256_u256(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_sub_cipheru256_u256, input) }
    pub fn sub_u256_cipheru256(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_sub_u256_cipheru256, input) }
    pub fn mul_cipheru256_cipheru256(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_mul_cipheru256_cipheru256, input) }
    pub fn mul_cipheru256_u256(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_mul_cipheru256_u256, input) }
    pub fn mul_u256_cipheru256(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_mul_u256_cipheru256, input) }
    pub fn add_cipheru64_cipheru64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_add_cipheru64_cipheru64, input) }
    pub fn add_cipheru64_u64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_add_cipheru64_u64, input) }
    pub fn add_u64_cipheru64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_add_u64_cipheru64, input) }
    pub fn sub_cipheru64_cipheru64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_sub_cipheru64_cipheru64, input) }
    pub fn sub_cipheru64_u64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_sub_cipheru64_u64, input) }
    pub fn sub_u64_cipheru64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_sub_u64_cipheru64, input) }
    pub fn mul_cipheru64_cipheru64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_mul_cipheru64_cipheru64, input) }
    pub fn mul_cipheru64_u64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_mul_cipheru64_u64, input) }
    pub fn mul_u64_cipheru64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_mul_u64_cipheru64, input) }
    pub fn add_cipheri64_cipheri64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_add_cipheri64_cipheri64, input) }
    pub fn add_cipheri64_i64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_add_cipheri64_i64, input) }
    pub fn add_i64_cipheri64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_add_i64_cipheri64, input) }
    pub fn sub_cipheri64_cipheri64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_sub_cipheri64_cipheri64, input) }
    pub fn sub_cipheri64_i64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_sub_cipheri64_i64, input) }
    pub fn sub_i64_cipheri64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_sub_i64_cipheri64, input) }
    pub fn mul_cipheri64_cipheri64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_mul_cipheri64_cipheri64, input) }
    pub fn mul_cipheri64_i64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_mul_cipheri64_i64, input) }
    pub fn mul_i64_cipheri64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_mul_i64_cipheri64, input) }
    pub fn add_cipherfrac64_cipherfrac64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_add_cipherfrac64_cipherfrac64, input) }
    pub fn add_cipherfrac64_frac64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_add_cipherfrac64_frac64, input) }
    pub fn add_frac64_cipherfrac64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_add_frac64_cipherfrac64, input) }
    pub fn sub_cipherfrac64_cipherfrac64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_sub_cipherfrac64_cipherfrac64, input) }
    pub fn sub_cipherfrac64_frac64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_sub_cipherfrac64_frac64, input) }
    pub fn sub_frac64_cipherfrac64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_sub_frac64_cipherfrac64, input) }
    pub fn mul_cipherfrac64_cipherfrac64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_mul_cipherfrac64_cipherfrac64, input) }
    pub fn mul_cipherfrac64_frac64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_mul_cipherfrac64_frac64, input) }
    pub fn mul_frac64_cipherfrac64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_mul_frac64_cipherfrac64, input) }
    pub fn encrypt_u256(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_encrypt_u256, input) }
    pub fn encrypt_u64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_encrypt_u64, input) }
    pub fn encrypt_i64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_encrypt_i64, input) }
    pub fn encrypt_frac64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_encrypt_frac64, input) }
    pub fn reencrypt_u256(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_reencrypt_u256, input) }
    pub fn reencrypt_u64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_reencrypt_u64, input) }
    pub fn reencrypt_i64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_reencrypt_i64, input) }
    pub fn reencrypt_frac64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_reencrypt_frac64, input) }
    pub fn decrypt_u256(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_decrypt_u256, input) }
    pub fn decrypt_u64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_decrypt_u64, input) }
    pub fn decrypt_i64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_decrypt_i64, input) }
    pub fn decrypt_frac64(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_decrypt_frac64, input) }
    pub fn public_key_bytes(&self, input: &[u8]) -> PrecompileResult { call(c_fhe_public_key_bytes, input) }
}

pub static FHE: FheApp = FheApp;
