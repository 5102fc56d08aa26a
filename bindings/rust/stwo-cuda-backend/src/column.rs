//! `Column<T>`, `ColumnOps<T>` and `FieldOps<F>` (upstream `core/backend/mod.rs`) over `sc_col` handles.
//!
//! A `sc_col` is a device buffer of 32-bit words.  `BaseField` columns hold one word per element, `Blake2sHash` columns
//! eight; a `SecureField` column is four coordinate columns, the layout `SecureColumnByCoords` already uses.
//! `at` / `set` are single-element host↔device copies exactly as upstream's `Column::at` on a device backend would be;
//! the prover's own decommitment path avoids them through `sc_gather` (see `merkle.rs`).

use std::ptr;

use stwo_prover::core::backend::{Column, ColumnOps};
use stwo_prover::core::fields::m31::BaseField;
use stwo_prover::core::fields::qm31::SecureField;
use stwo_prover::core::fields::FieldOps;
use stwo_prover::core::vcs::blake2_hash::Blake2sHash;

use crate::{ck, ctx, ffi, words, CudaBackend};

/// Owning handle; `Drop` → `sc_col_free` (exactly once, as the ABI requires).
#[derive(Debug)]
pub struct RawCol(pub(crate) *mut ffi::ScCol);

impl RawCol {
    pub(crate) fn words(&self) -> u64 {
        unsafe { ffi::sc_col_len(self.0) }
    }
    pub(crate) fn zeros(words: u64) -> Self {
        let mut h = ptr::null_mut();
        ck(unsafe { ffi::sc_col_zeros(ctx(), words, &mut h) });
        RawCol(h)
    }
    pub(crate) fn uninit(words: u64) -> Self {
        let mut h = ptr::null_mut();
        ck(unsafe { ffi::sc_col_uninit(ctx(), words, &mut h) });
        RawCol(h)
    }
    pub(crate) fn from_host(host: &[u32]) -> Self {
        let mut h = ptr::null_mut();
        ck(unsafe { ffi::sc_col_from_host(ctx(), host.as_ptr(), host.len() as u64, &mut h) });
        RawCol(h)
    }
    pub(crate) fn to_host(&self) -> Vec<u32> {
        let mut v = vec![0u32; self.words() as usize];
        ck(unsafe { ffi::sc_col_to_host(ctx(), self.0, v.as_mut_ptr()) });
        v
    }
    pub(crate) fn read(&self, offset: u64, out: &mut [u32]) {
        ck(unsafe { ffi::sc_col_read(ctx(), self.0, offset, out.len() as u64, out.as_mut_ptr()) });
    }
    pub(crate) fn write(&mut self, offset: u64, src: &[u32]) {
        ck(unsafe { ffi::sc_col_write(ctx(), self.0, offset, src.len() as u64, src.as_ptr()) });
    }
}
impl Clone for RawCol {
    fn clone(&self) -> Self {
        let mut h = ptr::null_mut();
        ck(unsafe { ffi::sc_col_clone(ctx(), self.0, &mut h) });
        RawCol(h)
    }
}
impl Drop for RawCol {
    fn drop(&mut self) {
        if !self.0.is_null() {
            // a failing free must not panic inside an unwinding drop; the context reports it on the next call
            unsafe { ffi::sc_col_free(ctx(), self.0) };
        }
    }
}

// ------------------------------------------------------------------------------------------------ BaseField

#[derive(Clone, Debug)]
pub struct CudaBaseColumn(pub(crate) RawCol);

impl CudaBaseColumn {
    pub fn handle(&self) -> *mut ffi::ScCol {
        self.0 .0
    }
    pub(crate) fn from_handle(h: *mut ffi::ScCol) -> Self {
        CudaBaseColumn(RawCol(h))
    }
}

impl Column<BaseField> for CudaBaseColumn {
    fn zeros(len: usize) -> Self {
        CudaBaseColumn(RawCol::zeros(len as u64))
    }
    unsafe fn uninitialized(len: usize) -> Self {
        CudaBaseColumn(RawCol::uninit(len as u64))
    }
    fn to_cpu(&self) -> Vec<BaseField> {
        self.0.to_host().into_iter().map(BaseField::from_u32_unchecked).collect()
    }
    fn len(&self) -> usize {
        self.0.words() as usize
    }
    fn at(&self, index: usize) -> BaseField {
        let mut w = [0u32];
        self.0.read(index as u64, &mut w);
        BaseField::from_u32_unchecked(w[0])
    }
    fn set(&mut self, index: usize, value: BaseField) {
        self.0.write(index as u64, &[value.0]);
    }
}
impl FromIterator<BaseField> for CudaBaseColumn {
    fn from_iter<I: IntoIterator<Item = BaseField>>(iter: I) -> Self {
        let host: Vec<u32> = iter.into_iter().map(|v| v.0).collect();
        CudaBaseColumn(RawCol::from_host(&host))
    }
}

impl ColumnOps<BaseField> for CudaBackend {
    type Column = CudaBaseColumn;
    /// upstream `core/backend/simd/bit_reverse.rs`
    fn bit_reverse_column(column: &mut Self::Column) {
        ck(unsafe { ffi::sc_bit_reverse(ctx(), column.handle()) });
    }
}

impl FieldOps<BaseField> for CudaBackend {
    fn batch_inverse(column: &Self::Column, dst: &mut Self::Column) {
        ck(unsafe { ffi::sc_batch_inverse_m31(ctx(), column.handle(), dst.handle()) });
    }
}

// ------------------------------------------------------------------------------------------------ SecureField

/// Four coordinate columns.  Only `FieldOps<SecureField>` (a `Backend` super-trait) needs this type: the prover's
/// secure data lives in `SecureColumnByCoords<CudaBackend>`, i.e. four `CudaBaseColumn`s.
#[derive(Clone, Debug)]
pub struct CudaSecureColumn(pub(crate) [CudaBaseColumn; 4]);

impl CudaSecureColumn {
    fn handles(&self) -> [*mut ffi::ScCol; 4] {
        [self.0[0].handle(), self.0[1].handle(), self.0[2].handle(), self.0[3].handle()]
    }
}

impl Column<SecureField> for CudaSecureColumn {
    fn zeros(len: usize) -> Self {
        CudaSecureColumn(std::array::from_fn(|_| CudaBaseColumn::zeros(len)))
    }
    unsafe fn uninitialized(len: usize) -> Self {
        CudaSecureColumn(std::array::from_fn(|_| unsafe { CudaBaseColumn::uninitialized(len) }))
    }
    fn to_cpu(&self) -> Vec<SecureField> {
        let c: [Vec<u32>; 4] = std::array::from_fn(|k| self.0[k].0.to_host());
        (0..c[0].len()).map(|i| words::to_qm31(&[c[0][i], c[1][i], c[2][i], c[3][i]])).collect()
    }
    fn len(&self) -> usize {
        self.0[0].len()
    }
    fn at(&self, index: usize) -> SecureField {
        // one gather for the four coordinates instead of four copies
        let (h, off) = (self.handles(), [index as u64; 4]);
        let mut w = [0u32; 4];
        ck(unsafe { ffi::sc_gather(ctx(), h.as_ptr(), off.as_ptr(), 4, 1, w.as_mut_ptr()) });
        words::to_qm31(&w)
    }
    fn set(&mut self, index: usize, value: SecureField) {
        for (k, w) in words::qm31(value).into_iter().enumerate() {
            self.0[k].0.write(index as u64, &[w]);
        }
    }
}
impl FromIterator<SecureField> for CudaSecureColumn {
    fn from_iter<I: IntoIterator<Item = SecureField>>(iter: I) -> Self {
        let mut c: [Vec<u32>; 4] = Default::default();
        for v in iter {
            for (k, w) in words::qm31(v).into_iter().enumerate() {
                c[k].push(w);
            }
        }
        CudaSecureColumn(std::array::from_fn(|k| CudaBaseColumn(RawCol::from_host(&c[k]))))
    }
}

impl ColumnOps<SecureField> for CudaBackend {
    type Column = CudaSecureColumn;
    fn bit_reverse_column(column: &mut Self::Column) {
        for c in &column.0 {
            ck(unsafe { ffi::sc_bit_reverse(ctx(), c.handle()) });
        }
    }
}

impl FieldOps<SecureField> for CudaBackend {
    /// `LogupColGenerator::finalize_col` upstream; the reference reaches it from every `interaction_trace_evaluation`
    /// (e.g. components/memory/table.rs:513).  With `CudaLogup` that whole function is one call, so this method is
    /// kept for completeness of the trait family.
    fn batch_inverse(column: &Self::Column, dst: &mut Self::Column) {
        let (s, d) = (column.handles(), dst.handles());
        ck(unsafe { ffi::sc_batch_inverse_qm31(ctx(), s.as_ptr(), d.as_ptr()) });
    }
}

// ------------------------------------------------------------------------------------------------ Blake2sHash

/// Merkle layer: eight words per node.
#[derive(Clone, Debug)]
pub struct CudaHashColumn(pub(crate) RawCol);

impl CudaHashColumn {
    pub fn handle(&self) -> *mut ffi::ScCol {
        self.0 .0
    }
    pub(crate) fn from_handle(h: *mut ffi::ScCol) -> Self {
        CudaHashColumn(RawCol(h))
    }
}

fn hash_from_words(w: &[u32]) -> Blake2sHash {
    let mut b = [0u8; 32];
    for (k, x) in w.iter().enumerate() {
        b[4 * k..4 * k + 4].copy_from_slice(&x.to_le_bytes());
    }
    Blake2sHash(b)
}
fn hash_to_words(h: &Blake2sHash) -> [u32; 8] {
    std::array::from_fn(|k| u32::from_le_bytes(h.0[4 * k..4 * k + 4].try_into().unwrap()))
}

impl Column<Blake2sHash> for CudaHashColumn {
    fn zeros(len: usize) -> Self {
        CudaHashColumn(RawCol::zeros(8 * len as u64))
    }
    unsafe fn uninitialized(len: usize) -> Self {
        CudaHashColumn(RawCol::uninit(8 * len as u64))
    }
    fn to_cpu(&self) -> Vec<Blake2sHash> {
        self.0.to_host().chunks_exact(8).map(hash_from_words).collect()
    }
    fn len(&self) -> usize {
        (self.0.words() / 8) as usize
    }
    fn at(&self, index: usize) -> Blake2sHash {
        let mut w = [0u32; 8];
        self.0.read(8 * index as u64, &mut w);
        hash_from_words(&w)
    }
    fn set(&mut self, index: usize, value: Blake2sHash) {
        self.0.write(8 * index as u64, &hash_to_words(&value));
    }
}
impl FromIterator<Blake2sHash> for CudaHashColumn {
    fn from_iter<I: IntoIterator<Item = Blake2sHash>>(iter: I) -> Self {
        let host: Vec<u32> = iter.into_iter().flat_map(|h| hash_to_words(&h)).collect();
        CudaHashColumn(RawCol::from_host(&host))
    }
}

impl ColumnOps<Blake2sHash> for CudaBackend {
    type Column = CudaHashColumn;
    fn bit_reverse_column(_column: &mut Self::Column) {
        // upstream leaves this unimplemented for hash columns on every backend; nothing in the prover calls it
        unimplemented!("bit_reverse_column on a hash column")
    }
}
