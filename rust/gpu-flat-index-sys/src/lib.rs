//! Raw declarations of every entry point of `include/gfi.h` (what bindgen emits for it), in header order.
//! Semantics, ownership and the reference interface each call replaces are documented in the header.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

/// Opaque handle (`typedef struct gfi_index gfi_index`).
#[repr(C)]
pub struct gfi_index {
    _private: [u8; 0],
}

pub const GFI_OK: i32 = 0;
pub const GFI_ERR_DIMENSION_MISMATCH: i32 = 1;
pub const GFI_ERR_INVALID_VECTOR: i32 = 2;
pub const GFI_ERR_INDEX: i32 = 3;
pub const GFI_ERR_NAN: i32 = 4;
pub const GFI_ERR_UNPROVEN: i32 = 5;
pub const GFI_METRIC_EUCLIDEAN: i32 = 0;
pub const GFI_METRIC_COSINE: i32 = 1;
pub const GFI_METRIC_DOT: i32 = 2;
pub const GFI_FLAG_NO_TENSOR: u32 = 1;
pub const GFI_GEN_UNIFORM: i32 = 0;
pub const GFI_GEN_NORMAL: i32 = 1;

/// `gfi_stats` (all fields `int64_t`, same order as the header).
#[repr(C)]
#[derive(Debug, Default, Clone, Copy)]
pub struct gfi_stats {
    pub n_slots: i64,
    pub n_live: i64,
    pub searches: i64,
    pub queries: i64,
    pub scan_queries: i64,
    pub tensor_queries: i64,
    pub fallback_queries: i64,
    pub kernel_launches: i64,
    pub bytes_fp32: i64,
    pub bytes_fp16: i64,
    pub scan_kernel_ns: i64,
    pub scan_kernel_count: i64,
    pub tensor_kernel_ns: i64,
    pub tensor_kernel_count: i64,
    pub coalesced_batches: i64,
    pub coalesced_requests: i64,
    pub shards: i64,
    pub merge_ns: i64,
    pub merge_count: i64,
    pub paged_queries: i64,
}

extern "C" {
    pub fn gfi_create(out: *mut *mut gfi_index, metric: i32, dim: i64, device: i32, flags: u32) -> i32;
    pub fn gfi_destroy(h: *mut gfi_index) -> i32;
    pub fn gfi_create_sharded(out: *mut *mut gfi_index, metric: i32, dim: i64, devices: *const i32, n_devices: i32,
                              flags: u32) -> i32;
    pub fn gfi_add(h: *mut gfi_index, ids: *const u64, rows: *const f32, n: i64, dim: i64) -> i32;
    pub fn gfi_add_generated(h: *mut gfi_index, seed: u32, first_row: u64, n: i64, kind: i32, first_id: u64) -> i32;
    pub fn gfi_add_from_file(h: *mut gfi_index, path: *const c_char, first_id: u64, out_rows: *mut i64) -> i32;
    pub fn gfi_remove(h: *mut gfi_index, id: u64) -> i32;
    pub fn gfi_len(h: *const gfi_index) -> i64;
    pub fn gfi_metric(h: *const gfi_index) -> i32;
    pub fn gfi_dim(h: *const gfi_index) -> i64;
    pub fn gfi_get_vector(h: *mut gfi_index, id: u64, out: *mut f32, cap: i64, out_dim: *mut i64) -> i32;
    pub fn gfi_flush(h: *mut gfi_index) -> i32;
    pub fn gfi_reserve(h: *mut gfi_index, n_rows: i64) -> i32;
    pub fn gfi_compact(h: *mut gfi_index) -> i32;
    pub fn gfi_search(h: *mut gfi_index, queries: *const f32, q: i64, dim: i64, ks: *const u32, mask: *const u64,
                      mask_bits: i64, out_ids: *mut u64, out_dist: *mut f32, out_counts: *mut u32, kstride: i64) -> i32;
    pub fn gfi_set_metadata(h: *mut gfi_index, id: u64, n_fields: i32, keys: *const *const c_char,
                            values: *const *const c_char) -> i32;
    pub fn gfi_set_metadata_column(h: *mut gfi_index, key: *const c_char, ids: *const u64, n: i64,
                                   values: *const *const c_char, n_values: i32, codes: *const u32) -> i32;
    pub fn gfi_search_filtered(h: *mut gfi_index, queries: *const f32, q: i64, dim: i64, ks: *const u32,
                               filter_json: *const c_char, out_ids: *mut u64, out_dist: *mut f32,
                               out_counts: *mut u32, kstride: i64) -> i32;
    pub fn gfi_search_device(h: *mut gfi_index, d_queries: *const f32, q: i64, d_ks: *const u32, kmax: u32,
                             d_mask: *const u64, mask_bits: i64, d_out_ids: *mut u64, d_out_dist: *mut f32,
                             d_out_counts: *mut u32, kstride: i64, stream: *mut c_void) -> i32;
    pub fn gfi_search_status(h: *mut gfi_index) -> i32;
    pub fn gfi_merge_topk_device(d_ids: *const u64, d_dist: *const f32, d_counts: *const u32, g: i32, q: i64,
                                 kstride: i64, d_ks: *const u32, d_out_ids: *mut u64, d_out_dist: *mut f32,
                                 d_out_counts: *mut u32, out_kstride: i64, stream: *mut c_void) -> i32;
    pub fn gfi_merge_topk_device_strided(d_ids: *const u64, d_dist: *const f32, d_counts: *const u32, g: i32, q: i64,
                                         kstride: i64, shard_stride_bytes: i64, d_ks: *const u32,
                                         d_out_ids: *mut u64, d_out_dist: *mut f32, d_out_counts: *mut u32,
                                         out_kstride: i64, stream: *mut c_void) -> i32;
    pub fn gfi_distances(h: *mut gfi_index, queries: *const f32, q: i64, dim: i64, cand_ids: *const u64, m: i64,
                         out_dist: *mut f32, out_status: *mut u8) -> i32;
    pub fn gfi_last_mismatch(expected: *mut i64, actual: *mut i64);
    pub fn gfi_last_error() -> *const c_char;
    pub fn gfi_get_stats(h: *mut gfi_index, out: *mut gfi_stats) -> i32;
    pub fn gfi_debug_tensor_scores(h: *mut gfi_index, queries: *const f32, q: i64, out: *mut f32, out_stride: i64) -> i32;
    pub fn gfi_set_option(h: *mut gfi_index, name: *const c_char, value: i64) -> i32;
    pub fn gfi_version() -> i32;
}
