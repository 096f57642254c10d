//! Drop-in replacement for quaru's `src/register.rs`: same public items and method signatures
//! (reference: src/register.rs:10-20 `OperationError`, :23-37 `Register`, :39-407 `impl Register`,
//! :409-413 `PartialEq`), with the state held in B200 HBM behind the C ABI of include/quaru_b200.h.
//!
//! The one intentional API change (north_star): the public `state` FIELD cannot survive a
//! device-resident buffer, so it becomes `state()` (download) and `set_state()` (upload).
//! `operation.rs`, `math.rs` and `openqasm.rs` of the reference stay as they are.
//!
//! UNVERIFIED: this file cannot be compiled in the repository's build image (no cargo/rustc).  The same
//! entry points, error mapping and panic messages are verified through the Python mirror
//! (quaru_b200/register.py) by tests/.
use crate::ffi::*;
use crate::{
    math::{self, c64},
    operation::{Operation, QuantumCircuit, QuantumOperation},
};
use ndarray::Array2;
use rand::prelude::*;
use std::ffi::CStr;

/// Errors which can occur when an operation is applied on the register (register.rs:10-20).
#[derive(Debug, PartialEq, Eq)]
pub enum OperationError {
    InvalidTarget(usize),
    InvalidDimensions(usize, usize),
    InvalidArity(usize),
    InvalidQubit(usize),
}

/// A quantum register containing N qubits; the 2^N amplitudes live on the GPU.
#[derive(Debug)]
pub struct Register {
    handle: *mut qb_reg,
    size: usize,
}

// The handle is a plain pointer to library-owned memory; `&mut self` on every mutating method gives the
// one-thread-at-a-time guarantee the library needs.
unsafe impl Send for Register {}

fn check(rc: i32) -> Result<(), OperationError> {
    if rc == QB_OK {
        return Ok(());
    }
    let payload = unsafe { qb_last_error_payload() } as usize;
    match rc {
        QB_ERR_INVALID_TARGET => Err(OperationError::InvalidTarget(payload)),
        QB_ERR_INVALID_DIMENSIONS => Err(OperationError::InvalidDimensions(payload, payload)),
        QB_ERR_INVALID_ARITY => Err(OperationError::InvalidArity(payload)),
        QB_ERR_INVALID_QUBIT => Err(OperationError::InvalidQubit(payload)),
        _ => {
            // CUDA / NCCL / OOM / index out of bounds: there is no CPU fallback and no reference error for these
            let msg = unsafe { CStr::from_ptr(qb_last_error()) }.to_string_lossy().into_owned();
            panic!("quaru_b200: {msg}");
        }
    }
}

fn last_error_message() -> String {
    unsafe { CStr::from_ptr(qb_last_error()) }.to_string_lossy().into_owned()
}

impl Register {
    /// register.rs:41
    pub fn new(input_bits: &[bool]) -> Self {
        let k = input_bits.iter().enumerate().fold(0usize, |acc, (q, &b)| acc | ((b as usize) << q));
        Self::from_int(input_bits.len(), k)
    }

    /// register.rs:58
    pub fn from_int(n: usize, k: usize) -> Self {
        let mut handle = std::ptr::null_mut();
        check(unsafe { qb_create(n as u32, k as u64, &mut handle) }).expect("Could not create register");
        Self { handle, size: n }
    }

    /// register.rs:74
    pub fn from_qubits(input_bits: &[Array2<math::c64>]) -> Self {
        Self::try_from_qubits(input_bits).expect("Incorrect input qubits")
    }

    /// register.rs:81
    pub fn try_from_qubits(input_bits: &[Array2<math::c64>]) -> Result<Self, OperationError> {
        let mut flat: Vec<f64> = Vec::with_capacity(4 * input_bits.len());
        for (i, q) in input_bits.iter().enumerate() {
            if q.len() != 2 {
                return Err(OperationError::InvalidQubit(i)); // is_qubit (register.rs:120-126)
            }
            for v in q.iter() {
                flat.push(v.re);
                flat.push(v.im);
            }
        }
        let mut handle = std::ptr::null_mut();
        check(unsafe { qb_create_product(flat.as_ptr(), input_bits.len() as u32, &mut handle) })?;
        Ok(Self { handle, size: input_bits.len() })
    }

    /// register.rs:134
    pub fn apply(&mut self, op: &Operation) -> &mut Self {
        self.try_apply_fast(op).expect("Could not apply operation")
    }

    /// register.rs:140
    pub fn try_apply_fast(&mut self, op: &Operation) -> Result<&mut Self, OperationError> {
        let matrix = op.matrix(); // Array2<c64>, standard (row-major) layout; c64 is #[repr(C)] {re, im}
        let matrix = matrix.as_standard_layout();
        // a target that does not fit the ABI's u32 can never be a qubit of this register: report it as the reference would
        // (InvalidTarget(t), register.rs:149-153) instead of truncating it in the cast below
        if let Some(&t) = op.targets().iter().find(|&&t| t > u32::MAX as usize) {
            return Err(OperationError::InvalidTarget(t));
        }
        let targets: Vec<u32> = op.targets().iter().map(|&t| t as u32).collect();
        check(unsafe {
            qb_apply(self.handle, matrix.as_ptr() as *const f64, targets.as_ptr(), targets.len() as u32)
        })?;
        Ok(self)
    }

    /// register.rs:193 -- same contract as try_apply_fast
    pub fn try_apply(&mut self, op: &Operation) -> Result<&mut Self, OperationError> {
        self.try_apply_fast(op)
    }

    /// register.rs:259
    pub fn apply_all(&mut self, operation: &Operation) -> &mut Self {
        self.try_apply_all(operation).expect("Could not apply operation")
    }

    /// register.rs:267
    pub fn try_apply_all(&mut self, operation: &Operation) -> Result<&mut Self, OperationError> {
        if operation.arity() != 1 {
            return Err(OperationError::InvalidArity(operation.arity()));
        }
        let m = operation.matrix();
        let m = m.as_standard_layout();
        check(unsafe { qb_apply_all(self.handle, m.as_ptr() as *const f64) })?;
        Ok(self)
    }

    /// register.rs:291
    pub fn measure(&mut self, target: usize) -> bool {
        self.try_measure(target).expect("Could not measure bit")
    }

    /// register.rs:297 -- the uniform draw stays where the reference takes it (thread_rng, :317-318)
    pub fn try_measure(&mut self, target: usize) -> Result<bool, OperationError> {
        let x: f64 = rand::thread_rng().gen();
        self.try_measure_with(target, x)
    }

    /// Seeded-parity hook: measurement with a caller-supplied uniform draw.
    pub fn try_measure_with(&mut self, target: usize, u: f64) -> Result<bool, OperationError> {
        if target > u32::MAX as usize {
            return Err(OperationError::InvalidTarget(target)); // register.rs:298-300; never truncated by the cast below
        }
        let mut outcome: i32 = 0;
        check(unsafe { qb_measure(self.handle, target as u32, u, &mut outcome) })?;
        Ok(outcome != 0)
    }

    /// register.rs:345
    pub fn apply_circuit(&mut self, circuit: &mut QuantumCircuit) -> Vec<(usize, bool)> {
        // one C-ABI call for the whole gate list, so the library can fuse it into tile passes
        let mats: Vec<_> = circuit.get_operations().iter().map(|op| op.matrix().as_standard_layout().to_owned()).collect();
        if circuit.get_operations().iter().any(|op| op.targets().iter().any(|&t| t > u32::MAX as usize)) {
            panic!("Could not apply operation"); // what Register::apply does for an InvalidTarget (register.rs:136)
        }
        let tgts: Vec<Vec<u32>> =
            circuit.get_operations().iter().map(|op| op.targets().iter().map(|&t| t as u32).collect()).collect();
        let gates: Vec<qb_gate> = mats
            .iter()
            .zip(tgts.iter())
            .map(|(m, t)| qb_gate { matrix: m.as_ptr() as *const f64, targets: t.as_ptr(), k: t.len() as u32 })
            .collect();
        check(unsafe { qb_apply_batch(self.handle, gates.as_ptr(), gates.len() as u64) }).expect("Could not apply operation");
        let mut measures: Vec<(usize, bool)> = Vec::new();
        if circuit.get_measurement_targets().len() > 0 {
            measures = self.measure_circuit(circuit);
        }
        let conditional: Vec<_> = circuit.get_conditional_operations().clone();
        for c_op in conditional.iter() {
            if c_op.1.iter().all(|(target, value)| measures.contains(&(*target, *value))) {
                self.apply(&c_op.0);
            }
        }
        measures
    }

    /// register.rs:371
    pub fn measure_circuit(&mut self, circuit: &mut QuantumCircuit) -> Vec<(usize, bool)> {
        let targets: Vec<usize> = circuit.get_measurement_targets().clone();
        targets.into_iter().map(|t| (t, self.measure(t))).collect()
    }

    /// Replaces reads of the reference's `pub state` field (register.rs:35): downloads the amplitudes.
    pub fn state(&self) -> Array2<c64> {
        let len = 1usize << self.size;
        let mut out = Array2::<c64>::zeros((len, 1));
        check(unsafe { qb_read_state(self.handle, 0, len as u64, out.as_mut_ptr() as *mut f64) }).expect("read_state");
        out
    }

    /// Replaces writes to `pub state` (e.g. examples/quantum_teleportation.rs:56-61).
    pub fn set_state(&mut self, state: &Array2<c64>) {
        let s = state.as_standard_layout();
        check(unsafe { qb_write_state(self.handle, 0, s.len() as u64, s.as_ptr() as *const f64) }).expect("write_state");
    }

    /// Streams the device-resident state to a raw state file (one shard per file, see include/quaru_b200.h).
    pub fn save_state(&self, path: &std::path::Path) -> Result<(), String> {
        let p = std::ffi::CString::new(path.to_string_lossy().as_bytes()).map_err(|e| e.to_string())?;
        match unsafe { qb_save_state(self.handle, p.as_ptr()) } {
            0 => Ok(()),
            _ => Err(last_error_message()),
        }
    }

    /// Restores a file written by `save_state` into this register (same qubit count).
    pub fn load_state(&mut self, path: &std::path::Path) -> Result<(), String> {
        let p = std::ffi::CString::new(path.to_string_lossy().as_bytes()).map_err(|e| e.to_string())?;
        match unsafe { qb_load_state(self.handle, p.as_ptr()) } {
            0 => Ok(()),
            _ => Err(last_error_message()),
        }
    }

    /// Streams the amplitudes to the host in chunks of 2^20 (a 33-qubit state is 137 GB: `state()` would need that much
    /// host memory) and hands each one, with its index, to `f`.
    fn for_each_amplitude(&self, mut f: impl FnMut(usize, c64)) {
        const CHUNK: u64 = 1 << 20;
        let total: u64 = 1u64 << self.size;
        let mut buf = vec![c64::new(0.0, 0.0); CHUNK.min(total) as usize];
        let mut offset = 0u64;
        while offset < total {
            let count = CHUNK.min(total - offset);
            check(unsafe { qb_read_state(self.handle, offset, count, buf.as_mut_ptr() as *mut f64) }).expect("Could not read state");
            for (k, s) in buf[..count as usize].iter().enumerate() {
                f(offset as usize + k, *s);
            }
            offset += count;
        }
    }

    /// register.rs:379
    pub fn print_probabilities(&self) {
        let n = self.size;
        self.for_each_amplitude(|i, s| println!("{:0n$b}: {}%", i, s.norm_sqr() * 100.0));
    }

    /// register.rs:386
    pub fn print_nonzero_probabilities(&self) {
        let n = self.size;
        self.for_each_amplitude(|i, s| {
            if s.norm_sqr() > 0.0 {
                println!("{:0n$b}: {}%", i, s.norm_sqr() * 100.0);
            }
        });
    }

    /// register.rs:396
    pub fn print_state(&self) {
        let n = self.size;
        self.for_each_amplitude(|i, s| println!("{i:0n$b}: {s}"));
    }

    /// register.rs:404
    pub fn size(&self) -> usize {
        self.size
    }
}

impl Clone for Register {
    fn clone(&self) -> Self {
        let mut handle = std::ptr::null_mut();
        check(unsafe { qb_clone(self.handle, &mut handle) }).expect("Could not clone register");
        Self { handle, size: self.size }
    }
}

impl Drop for Register {
    fn drop(&mut self) {
        unsafe { qb_destroy(self.handle) }
    }
}

/// register.rs:409-413: all |a_i - b_i| < 1e-8, evaluated on the device.
impl PartialEq for Register {
    fn eq(&self, other: &Self) -> bool {
        if self.size != other.size {
            return false;
        }
        let mut d = 0.0f64;
        check(unsafe { qb_max_abs_diff(self.handle, other.handle, &mut d) }).expect("max_abs_diff");
        d < 1e-8
    }
}
