//! `GpuFlatIndex` — a drop-in `impl Index` (reference `src/index.rs:11-35`) backed by libgfi's hand-written
//! sm_100a kernels.  Constructed like `HnswIndex` is today: `VectorStore::with_index(GpuFlatIndex::new(metric)?)`
//! (`src/storage.rs:118-127`, `src/server/mod.rs:39-40`).  Every existing caller -- `VectorStore::search`,
//! `search_with_filter`, the batch loops, the HTTP handlers, the CLI -- works unchanged on top of it.
use std::collections::HashMap;
use std::ffi::{CStr, CString};
use std::os::raw::c_char;

use gpu_flat_index_sys as sys;
use vectordb_from_scratch::{
    distance::DistanceMetric,
    error::{Result, VectorDbError},
    index::Index,
    storage::{Metadata, MetadataFilter},
    vector::Vector,
};

#[derive(Debug)]
pub struct GpuFlatIndex {
    h: *mut sys::gfi_index,
    metric: DistanceMetric,
    /// `Index::get_vector` lends `&Vector` (src/index.rs:23): the rows are mirrored on the host for that one call.
    mirror: HashMap<usize, Vector>,
}

// libgfi: searches are re-entrant, mutations exclusive -- the RwLock discipline of src/server/mod.rs:13-16.
unsafe impl Send for GpuFlatIndex {}
unsafe impl Sync for GpuFlatIndex {}

fn last_error() -> String {
    unsafe { CStr::from_ptr(sys::gfi_last_error()).to_string_lossy().into_owned() }
}

fn status(rc: i32) -> Result<()> {
    match rc {
        sys::GFI_OK => Ok(()),
        sys::GFI_ERR_DIMENSION_MISMATCH => {
            let (mut e, mut a) = (0i64, 0i64);
            unsafe { sys::gfi_last_mismatch(&mut e, &mut a) };
            Err(VectorDbError::DimensionMismatch { expected: e as usize, actual: a as usize })
        }
        sys::GFI_ERR_INVALID_VECTOR => Err(VectorDbError::InvalidVector {
            reason: "Cannot compute cosine distance with zero vector".to_string(),
        }),
        // GFI_ERR_NAN: the reference panics in sort_by (flat_index.rs:62); reported as an IndexError instead
        _ => Err(VectorDbError::IndexError(last_error())),
    }
}

fn metric_code(metric: DistanceMetric) -> i32 {
    match metric {
        DistanceMetric::Euclidean => sys::GFI_METRIC_EUCLIDEAN,
        DistanceMetric::Cosine => sys::GFI_METRIC_COSINE,
        DistanceMetric::DotProduct => sys::GFI_METRIC_DOT,
    }
}

impl GpuFlatIndex {
    /// One GPU (device 0).  The dimension is latched by the first `add`, as in `VectorStore`.
    pub fn new(metric: DistanceMetric) -> Result<Self> {
        Self::on_device(metric, 0)
    }

    pub fn on_device(metric: DistanceMetric, device: i32) -> Result<Self> {
        let mut h = std::ptr::null_mut();
        status(unsafe { sys::gfi_create(&mut h, metric_code(metric), 0, device, 0) })?;
        Ok(Self { h, metric, mirror: HashMap::new() })
    }

    /// ONE index sharded row-wise over several GPUs of the box (`devices[0]` merges).  The server owns a single
    /// `RwLock<VectorStore<I>>` (src/server/mod.rs:13-16), so this is how it reaches all eight GPUs.
    pub fn new_sharded(metric: DistanceMetric, devices: &[i32]) -> Result<Self> {
        let mut h = std::ptr::null_mut();
        status(unsafe {
            sys::gfi_create_sharded(&mut h, metric_code(metric), 0, devices.as_ptr(), devices.len() as i32, 0)
        })?;
        Ok(Self { h, metric, mirror: HashMap::new() })
    }

    /// Bulk load (one FFI call for n rows): `BatchInsertItem`s of `VectorStore::insert_batch` (src/storage.rs:190-215).
    pub fn add_batch(&mut self, ids: &[usize], rows: &[Vector]) -> Result<()> {
        if ids.is_empty() {
            return Ok(());
        }
        let dim = rows[0].dimension();
        let mut flat = Vec::with_capacity(ids.len() * dim);
        for r in rows {
            if r.dimension() != dim {
                return Err(VectorDbError::DimensionMismatch { expected: dim, actual: r.dimension() });
            }
            flat.extend_from_slice(r.as_slice());
        }
        let ids64: Vec<u64> = ids.iter().map(|&i| i as u64).collect();
        status(unsafe { sys::gfi_add(self.h, ids64.as_ptr(), flat.as_ptr(), ids.len() as i64, dim as i64) })?;
        for (i, r) in ids.iter().zip(rows) {
            self.mirror.insert(*i, r.clone());
        }
        Ok(())
    }

    /// Bulk load of the reference's flat vector file (src/persistence/mmap.rs:13-15,161-172).
    pub fn add_from_file(&mut self, path: &str, first_id: usize) -> Result<usize> {
        let p = CString::new(path).map_err(|e| VectorDbError::IndexError(e.to_string()))?;
        let mut n = 0i64;
        status(unsafe { sys::gfi_add_from_file(self.h, p.as_ptr(), first_id as u64, &mut n) })?;
        Ok(n as usize)
    }

    /// Hook for `VectorStore::insert_with_metadata`, right after `self.metadata.insert(internal_id, metadata)`
    /// (src/storage.rs:169): the fields of `id` are replaced and kept as dictionary-encoded columns in HBM.
    pub fn set_metadata(&mut self, id: usize, metadata: &Metadata) -> Result<()> {
        let keys: Vec<CString> = metadata.fields.keys().map(|k| CString::new(k.as_str()).unwrap()).collect();
        let vals: Vec<CString> = metadata.fields.values().map(|v| CString::new(v.as_str()).unwrap()).collect();
        let kp: Vec<*const c_char> = keys.iter().map(|k| k.as_ptr()).collect();
        let vp: Vec<*const c_char> = vals.iter().map(|v| v.as_ptr()).collect();
        status(unsafe { sys::gfi_set_metadata(self.h, id as u64, kp.len() as i32, kp.as_ptr(), vp.as_ptr()) })
    }

    /// `search_with_filter` with the filter handed down (exact pre-filter on the GPU) instead of the reference's
    /// post-filter over `fetch_k = 3k` (src/storage.rs:249-290).  `MetadataFilter` serialises to the JSON form
    /// libgfi parses (serde tag "op", src/storage.rs:44-58).
    pub fn search_filtered(&self, query: &Vector, k: usize, filter: &MetadataFilter) -> Result<Vec<(usize, f32)>> {
        let json = CString::new(serde_json::to_string(filter).map_err(|e| VectorDbError::IndexError(e.to_string()))?)
            .map_err(|e| VectorDbError::IndexError(e.to_string()))?;
        let ks = [k as u32];
        let cap = k.max(1);
        let (mut ids, mut dist, mut cnt) = (vec![0u64; cap], vec![0f32; cap], [0u32]);
        status(unsafe {
            sys::gfi_search_filtered(self.h, query.as_slice().as_ptr(), 1, query.dimension() as i64, ks.as_ptr(),
                                     json.as_ptr(), ids.as_mut_ptr(), dist.as_mut_ptr(), cnt.as_mut_ptr(), cap as i64)
        })?;
        Ok((0..cnt[0] as usize).map(|i| (ids[i] as usize, dist[i])).collect())
    }

    /// The defaulted trait extension `Index::search_batch`: ONE call for the whole batch
    /// with per-query k, as `VectorStore::search_batch` needs (src/storage.rs:302-310).  Fail-fast like the
    /// reference's `?` inside the loop: the first failing query fails the batch.
    pub fn search_batch(&self, queries: &[(Vector, usize)]) -> Result<Vec<Vec<(usize, f32)>>> {
        if queries.is_empty() {
            return Ok(Vec::new());
        }
        let dim = queries[0].0.dimension();
        if queries.iter().any(|(q, _)| q.dimension() != dim) {
            return queries.iter().map(|(q, k)| self.search(q, *k)).collect(); // per-query dimension errors
        }
        let q = queries.len();
        let kmax = queries.iter().map(|(_, k)| *k).max().unwrap_or(0).max(1);
        let mut flat = Vec::with_capacity(q * dim);
        for (v, _) in queries {
            flat.extend_from_slice(v.as_slice());
        }
        let ks: Vec<u32> = queries.iter().map(|(_, k)| *k as u32).collect();
        let (mut ids, mut dist, mut cnt) = (vec![0u64; q * kmax], vec![0f32; q * kmax], vec![0u32; q]);
        status(unsafe {
            sys::gfi_search(self.h, flat.as_ptr(), q as i64, dim as i64, ks.as_ptr(), std::ptr::null(), 0,
                            ids.as_mut_ptr(), dist.as_mut_ptr(), cnt.as_mut_ptr(), kmax as i64)
        })?;
        Ok((0..q)
            .map(|i| (0..cnt[i] as usize).map(|j| (ids[i * kmax + j] as usize, dist[i * kmax + j])).collect())
            .collect())
    }

    /// The defaulted trait extension `Index::search_masked`: eligibility by internal id as a bitmask.
    pub fn search_masked(&self, query: &Vector, k: usize, n_ids: usize, eligible: &dyn Fn(usize) -> bool)
                         -> Result<Vec<(usize, f32)>> {
        let mut words = vec![0u64; (n_ids + 63) / 64];
        for id in 0..n_ids {
            if eligible(id) {
                words[id / 64] |= 1u64 << (id % 64);
            }
        }
        let ks = [k as u32];
        let cap = k.max(1);
        let (mut ids, mut dist, mut cnt) = (vec![0u64; cap], vec![0f32; cap], [0u32]);
        status(unsafe {
            sys::gfi_search(self.h, query.as_slice().as_ptr(), 1, query.dimension() as i64, ks.as_ptr(), words.as_ptr(),
                            n_ids as i64, ids.as_mut_ptr(), dist.as_mut_ptr(), cnt.as_mut_ptr(), cap as i64)
        })?;
        Ok((0..cnt[0] as usize).map(|i| (ids[i] as usize, dist[i])).collect())
    }

    /// Exact `DistanceMetric::distance` of one vector against a list of stored rows in one call: the candidate
    /// evaluation of `HnswIndex::search_layer` (src/hnsw/graph.rs:221-232).  `None` where the id is absent or the
    /// reference would return `Err(InvalidVector)` (graph.rs: `unwrap_or(f32::MAX)`).
    pub fn distances(&self, v: &Vector, ids: &[usize]) -> Result<Vec<Option<f32>>> {
        let ids64: Vec<u64> = ids.iter().map(|&i| i as u64).collect();
        let (mut dist, mut st) = (vec![0f32; ids.len()], vec![0u8; ids.len()]);
        status(unsafe {
            sys::gfi_distances(self.h, v.as_slice().as_ptr(), 1, v.dimension() as i64, ids64.as_ptr(), ids.len() as i64,
                               dist.as_mut_ptr(), st.as_mut_ptr())
        })?;
        Ok(dist.iter().zip(&st).map(|(d, s)| if *s == 0 { Some(*d) } else { None }).collect())
    }

    pub fn stats(&self) -> Result<sys::gfi_stats> {
        let mut s = sys::gfi_stats::default();
        status(unsafe { sys::gfi_get_stats(self.h, &mut s) })?;
        Ok(s)
    }
}

impl Index for GpuFlatIndex {
    fn add(&mut self, id: usize, vector: Vector) -> Result<()> {
        // src/index.rs:13; FlatIndex::add is HashMap::insert (flat_index.rs:38-41): an existing id is overwritten
        let ids = [id as u64];
        status(unsafe { sys::gfi_add(self.h, ids.as_ptr(), vector.as_slice().as_ptr(), 1, vector.dimension() as i64) })?;
        self.mirror.insert(id, vector); // staged in pinned memory; flushed lazily by the next search
        Ok(())
    }

    fn remove(&mut self, id: usize) -> Result<()> {
        // src/index.rs:16; idempotent like HashMap::remove (flat_index.rs:44-47)
        self.mirror.remove(&id);
        status(unsafe { sys::gfi_remove(self.h, id as u64) })
    }

    fn search(&self, query: &Vector, k: usize) -> Result<Vec<(usize, f32)>> {
        // src/index.rs:20; flat_index.rs:52-65: score every row, sort ascending (lower id wins ties), truncate(k)
        let ks = [k as u32];
        let cap = k.max(1);
        let (mut ids, mut dist, mut cnt) = (vec![0u64; cap], vec![0f32; cap], [0u32]);
        status(unsafe {
            sys::gfi_search(self.h, query.as_slice().as_ptr(), 1, query.dimension() as i64, ks.as_ptr(),
                            std::ptr::null(), 0, ids.as_mut_ptr(), dist.as_mut_ptr(), cnt.as_mut_ptr(), cap as i64)
        })?;
        Ok((0..cnt[0] as usize).map(|i| (ids[i] as usize, dist[i])).collect())
    }

    fn get_vector(&self, id: usize) -> Option<&Vector> {
        self.mirror.get(&id) // src/index.rs:23
    }

    fn metric(&self) -> DistanceMetric {
        self.metric
    }

    fn len(&self) -> usize {
        unsafe { sys::gfi_len(self.h) as usize }
    }
}

impl Drop for GpuFlatIndex {
    fn drop(&mut self) {
        unsafe { sys::gfi_destroy(self.h) };
    }
}
