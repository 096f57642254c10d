//! Raw bindings of include/quaru_b200.h (what `bindgen` would emit), kept by hand: the ABI is small.
//! UNVERIFIED here (no rustc in the build image); the same symbols are exercised through ctypes by tests/.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_double, c_float, c_int, c_void};

#[repr(C)]
pub struct qb_reg {
    _private: [u8; 0],
}

#[repr(C)]
pub struct qb_gate {
    pub matrix: *const c_double,
    pub targets: *const u32,
    pub k: u32,
}

pub const QB_OK: c_int = 0;
pub const QB_ERR_INVALID_TARGET: c_int = -1;
pub const QB_ERR_INVALID_DIMENSIONS: c_int = -2;
pub const QB_ERR_INVALID_ARITY: c_int = -3;
pub const QB_ERR_INVALID_QUBIT: c_int = -4;
pub const QB_ERR_INVALID_ARGUMENT: c_int = -5;

extern "C" {
    pub fn qb_last_error() -> *const c_char;
    pub fn qb_last_error_payload() -> i64;
    pub fn qb_create(n: u32, basis_index: u64, out: *mut *mut qb_reg) -> c_int;
    pub fn qb_create_product(qubits: *const c_double, n: u32, out: *mut *mut qb_reg) -> c_int;
    pub fn qb_create_sharded(n: u32, basis_index: u64, rank: i32, world: i32, nccl_unique_id: *const c_void,
                             device: i32, out: *mut *mut qb_reg) -> c_int;
    pub fn qb_nccl_unique_id(out128: *mut c_void) -> c_int;
    pub fn qb_clone(src: *mut qb_reg, out: *mut *mut qb_reg) -> c_int;
    pub fn qb_destroy(reg: *mut qb_reg);
    pub fn qb_reset(reg: *mut qb_reg, basis_index: u64) -> c_int;
    pub fn qb_num_qubits(reg: *const qb_reg) -> u32;
    pub fn qb_apply(reg: *mut qb_reg, matrix: *const c_double, targets: *const u32, k: u32) -> c_int;
    pub fn qb_apply_controlled(reg: *mut qb_reg, matrix: *const c_double, targets: *const u32, k: u32,
                               controls: *const u32, n_controls: u32) -> c_int;
    pub fn qb_apply_all(reg: *mut qb_reg, m2x2: *const c_double) -> c_int;
    pub fn qb_apply_batch(reg: *mut qb_reg, gates: *const qb_gate, count: u64) -> c_int;
    pub fn qb_probabilities(reg: *mut qb_reg, target: u32, p0: *mut c_double, p1: *mut c_double) -> c_int;
    pub fn qb_measure(reg: *mut qb_reg, target: u32, u: c_double, outcome: *mut i32) -> c_int;
    pub fn qb_read_state(reg: *mut qb_reg, offset: u64, count: u64, out: *mut c_double) -> c_int;
    pub fn qb_write_state(reg: *mut qb_reg, offset: u64, count: u64, input: *const c_double) -> c_int;
    pub fn qb_save_state(reg: *mut qb_reg, path: *const c_char) -> c_int;
    pub fn qb_load_state(reg: *mut qb_reg, path: *const c_char) -> c_int;
    pub fn qb_norm2(reg: *mut qb_reg, out: *mut c_double) -> c_int;
    pub fn qb_max_abs_diff(a: *mut qb_reg, b: *mut qb_reg, out: *mut c_double) -> c_int;
    pub fn qb_l2_diff2(a: *mut qb_reg, b: *mut qb_reg, out: *mut c_double) -> c_int;
    pub fn qb_flush(reg: *mut qb_reg) -> c_int;
    pub fn qb_sync(reg: *mut qb_reg) -> c_int;
    pub fn qb_timer_start(reg: *mut qb_reg) -> c_int;
    pub fn qb_timer_stop(reg: *mut qb_reg, ms: *mut c_float) -> c_int;
    pub fn qb_set_option(reg: *mut qb_reg, option: i32, value: i64) -> c_int;
    pub fn qb_get_stat(reg: *mut qb_reg, stat: i32, value: *mut i64) -> c_int;
}
