//! jtk-gpu-sys -- binding of `libjtkgpu.so` (`include/jtk_gpu.h`) plus safe wrappers that carry the names and
//! argument meaning of the four kiley calls on jtk's per-chunk hot path, so that the reference's call sites change
//! by one receiver only:
//!
//! | kiley call (reference site)                                                           | wrapper here |
//! |---|---|
//! | `PairHiddenMarkovModel::modification_table_antidiagonal` (`pseudo_mcmc.rs:62-63`)      | [`GpuHmm::modification_table_antidiagonal`] |
//! | `PairHiddenMarkovModel::likelihood_antidiagonal_bootstrap` (`likelihood_gains.rs:27-28,282-283,301-302`) | [`GpuHmm::likelihood_antidiagonal_bootstrap`] |
//! | `PairHiddenMarkovModelOnStrands::polish_until_converge_antidiagonal` (`local_clustering/mod.rs:106,155-156`, `model_tune.rs:143`, `consensus/mod.rs:477-483`) | [`GpuHmm::polish_until_converge_antidiagonal`] |
//! | `PairHiddenMarkovModelOnStrands::fit_antidiagonal_par_multiple` (`model_tune.rs:145-151`) | [`GpuHmm::fit_antidiagonal_par_multiple`] |
//!
//! The batch-level entry points (`jtk_batch_*`, `jtk_polish_until_converge_batch` over many chunks) are what
//! `local_clustering_selected` should call for throughput; the single-call wrappers exist
//! for parity checks against kiley and for the low-volume call sites.
//!
//! This crate cannot be compiled in the development container (no Rust toolchain).  Every
//! declaration below is a hand transcription of `include/jtk_gpu.h`; `tests/test_abi_exports.py` checks that each
//! `#[repr(C)]` struct exists here and that the sizes pinned in the header hold in the compiled library.
#![allow(non_camel_case_types, clippy::too_many_arguments)]

use std::ffi::CStr;
use std::os::raw::{c_char, c_int};

pub const NUM_ROW: usize = 14; // kiley::hmm::NUM_ROW   (pseudo_mcmc.rs:7)
pub const COPY_SIZE: usize = 3; // kiley::hmm::COPY_SIZE (pseudo_mcmc.rs:172)
pub const DEL_SIZE: usize = 3;
pub const TABLE_NEG: f64 = -1.0e10;

/// `jtk_hmm_params` == `definitions::HMMParam` (definitions/src/lib.rs:102-126), 45 doubles = 360 bytes.
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct HmmParams {
    pub mat_mat: f64,
    pub mat_ins: f64,
    pub mat_del: f64,
    pub ins_mat: f64,
    pub ins_ins: f64,
    pub ins_del: f64,
    pub del_mat: f64,
    pub del_ins: f64,
    pub del_del: f64,
    pub mat_emit: [f64; 16],
    pub ins_emit: [f64; 20],
}
const _: () = assert!(std::mem::size_of::<HmmParams>() == 360);

/// `jtk_colstat`, 24 bytes.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct Colstat {
    pub sum: f64,
    pub count: i32,
    pub sc: [u16; 4],
    pub pad_: i32,
}
const _: () = assert!(std::mem::size_of::<Colstat>() == 24);

/// `jtk_candidate`, 32 bytes.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct Candidate {
    pub tmpl: u32,
    pub pos: u32,
    pub count: u32,
    pub pad_: u32,
    pub sum: f64,
    pub lk: f64,
}
const _: () = assert!(std::mem::size_of::<Candidate>() == 32);

/// `jtk_gains` == `likelihood_gains::Gains` flattened: rows Subst, Del, Ins x homopolymer length 1..=homop_len.
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct Gains {
    pub homop_len: c_int,
    pub gain: *const f64,
    pub prob: *const f64,
}
const _: () = assert!(std::mem::size_of::<Gains>() == 24);

/// `jtk_clustering_config` == `pseudo_mcmc::ClusteringConfig` without the borrowed gains (pseudo_mcmc.rs:18-43).
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct ClusteringConfig {
    pub band_width: c_int,
    pub copy_num: c_int,
    pub coverage: f64,
    pub local_coverage: f64,
}
const _: () = assert!(std::mem::size_of::<ClusteringConfig>() == 24);

/// `jtk_polish_config` == `kiley::hmm::HMMPolishConfig::new(radius, take_num, ignore_edge)`.
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct PolishConfig {
    pub radius: c_int,
    pub take_num: c_int,
    pub ignore_edge: c_int,
}
const _: () = assert!(std::mem::size_of::<PolishConfig>() == 12);

pub enum jtk_ctx {}
pub enum jtk_batch {}

extern "C" {
    pub fn jtk_ctx_create(device: c_int, workspace_bytes: usize, out: *mut *mut jtk_ctx) -> c_int;
    pub fn jtk_ctx_destroy(ctx: *mut jtk_ctx);
    pub fn jtk_last_error(ctx: *const jtk_ctx) -> *const c_char;
    pub fn jtk_hmm_num_row() -> c_int;
    pub fn jtk_hmm_copy_size() -> c_int;
    pub fn jtk_hmm_del_size() -> c_int;
    pub fn jtk_abi_sizeof(name: *const c_char) -> usize;

    pub fn jtk_hmm_modtable_batch(
        ctx: *mut jtk_ctx, fwd: *const HmmParams, rev: *const HmmParams, n_pairs: c_int, n_tmpl: c_int,
        tmpl_concat: *const u8, tmpl_off: *const u32, read_concat: *const u8, read_off: *const u32,
        ops_concat: *const u8, ops_off: *const u32, strand: *const u8, tmpl_idx: *const u32, radius: c_int,
        out_lk: *mut f64, out_table: *mut f64, table_off: *const u64,
    ) -> c_int;
    pub fn jtk_hmm_likelihood_batch(
        ctx: *mut jtk_ctx, fwd: *const HmmParams, rev: *const HmmParams, n_pairs: c_int, n_tmpl: c_int,
        tmpl_concat: *const u8, tmpl_off: *const u32, read_concat: *const u8, read_off: *const u32,
        ops_concat: *const u8, ops_off: *const u32, strand: *const u8, tmpl_idx: *const u32, radius: c_int,
        out_lk: *mut f64,
    ) -> c_int;

    pub fn jtk_batch_create(
        ctx: *mut jtk_ctx, n_pairs: c_int, n_tmpl: c_int, tmpl_concat: *const u8, tmpl_off: *const u32,
        read_concat: *const u8, read_off: *const u32, ops_concat: *const u8, ops_off: *const u32, strand: *const u8,
        tmpl_idx: *const u32, radius: c_int, out: *mut *mut jtk_batch,
    ) -> c_int;
    pub fn jtk_batch_destroy(b: *mut jtk_batch);
    pub fn jtk_batch_cell_updates(b: *const jtk_batch) -> u64;
    pub fn jtk_batch_modtable(b: *mut jtk_batch, fwd: *const HmmParams, rev: *const HmmParams, rows: c_int) -> c_int;
    pub fn jtk_batch_sync(b: *mut jtk_batch) -> c_int;
    pub fn jtk_batch_fetch_lk(b: *mut jtk_batch, out_lk: *mut f64) -> c_int;
    pub fn jtk_batch_fetch_profile(b: *mut jtk_batch, pair: c_int, out: *mut f32) -> c_int;
    pub fn jtk_batch_colstats(
        b: *mut jtk_batch, min_req: *const f32, h: c_int, pos_thr: f32, out: *mut Colstat, stat_off: *const u64,
    ) -> c_int;
    pub fn jtk_batch_gather(
        b: *mut jtk_batch, tmpl: c_int, min_req: *const f32, h: c_int, cols: *const u32, d: c_int, out: *mut f64,
    ) -> c_int;
    pub fn jtk_batch_candidates(
        b: *mut jtk_batch, gains: *const Gains, copy_num: *const i32, coverage: f64, out: *mut Candidate, cap: c_int,
        out_n: *mut c_int,
    ) -> c_int;
    pub fn jtk_batch_search_variants(
        b: *mut jtk_batch, gains: *const Gains, copy_num: *const i32, coverage: f64, probe_cap: c_int,
        out_n_probes: *mut u32, out_probe_pos: *mut u32, out_variants: *mut f64,
    ) -> c_int;

    pub fn jtk_polish_until_converge_batch(
        ctx: *mut jtk_ctx, fwd: *const HmmParams, rev: *const HmmParams, n_chunks: c_int, draft_concat: *const u8,
        draft_off: *const u32, n_pairs: c_int, read_concat: *const u8, read_off: *const u32, ops_buf: *mut u8,
        ops_pos: *const u64, ops_cap: *const u32, n_ops: *mut u32, strand: *const u8, tmpl_idx: *const u32,
        cfg: *const PolishConfig, out_cons: *mut u8, cons_pos: *const u64, cons_cap: *const u32, out_len: *mut u32,
        out_iters: *mut i32,
    ) -> c_int;
    pub fn jtk_hmm_fit_batch(
        ctx: *mut jtk_ctx, fwd: *mut HmmParams, rev: *mut HmmParams, n_pairs: c_int, n_tmpl: c_int,
        tmpl_concat: *const u8, tmpl_off: *const u32, read_concat: *const u8, read_off: *const u32,
        ops_concat: *const u8, ops_off: *const u32, strand: *const u8, tmpl_idx: *const u32, radius: c_int,
    ) -> c_int;

    pub fn jtk_mcmc_restarts_batch(
        ctx: *mut jtk_ctx, n_chains: c_int, data_concat: *const f64, data_off: *const u64, n_rows: *const u32,
        n_cols: *const u32, n_clusters: *const u32, size_to_lk_concat: *const f64, restarts: c_int,
        rng_state: *mut u64, out_asn: *mut u8, out_lk: *mut f64, out_err: *mut c_int,
    ) -> c_int;
}

/// One alignment column, byte-compatible with the `ops` arrays of the C ABI and in the order of `kiley::Op`
/// (`haplotyper/src/misc.rs:167-172`).
#[repr(u8)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Op {
    Match = 0,
    Mismatch = 1,
    Ins = 2,
    Del = 3,
}

#[cfg(feature = "kiley")]
impl From<kiley::Op> for Op {
    fn from(o: kiley::Op) -> Op {
        match o {
            kiley::Op::Match => Op::Match,
            kiley::Op::Mismatch => Op::Mismatch,
            kiley::Op::Ins => Op::Ins,
            kiley::Op::Del => Op::Del,
        }
    }
}
#[cfg(feature = "kiley")]
impl From<Op> for kiley::Op {
    fn from(o: Op) -> kiley::Op {
        match o {
            Op::Match => kiley::Op::Match,
            Op::Mismatch => kiley::Op::Mismatch,
            Op::Ins => kiley::Op::Ins,
            Op::Del => kiley::Op::Del,
        }
    }
}
/// `def_into_kiley` / `kiley_into_def` of `model_tune.rs:36-92`, to the ABI struct: every field of kiley's model is `pub`.
#[cfg(feature = "kiley")]
impl From<&kiley::hmm::PairHiddenMarkovModel> for HmmParams {
    fn from(m: &kiley::hmm::PairHiddenMarkovModel) -> HmmParams {
        HmmParams {
            mat_mat: m.mat_mat, mat_ins: m.mat_ins, mat_del: m.mat_del,
            ins_mat: m.ins_mat, ins_ins: m.ins_ins, ins_del: m.ins_del,
            del_mat: m.del_mat, del_ins: m.del_ins, del_del: m.del_del,
            mat_emit: m.mat_emit, ins_emit: m.ins_emit,
        }
    }
}

/// `TrainingDataPack::new(cons, strands, seqs, ops)` (`model_tune.rs:145-150`), borrowed.
pub struct TrainingDataPack<'a, T: AsRef<[u8]>> {
    pub consensus: &'a [u8],
    pub directions: &'a [bool],
    pub sequences: &'a [T],
    pub operations: &'a [Vec<Op>],
}
impl<'a, T: AsRef<[u8]>> TrainingDataPack<'a, T> {
    pub fn new(consensus: &'a [u8], directions: &'a [bool], sequences: &'a [T], operations: &'a [Vec<Op>]) -> Self {
        TrainingDataPack { consensus, directions, sequences, operations }
    }
}

/// One GPU context.  Calls on one context are serialised by `&mut self`; use one `GpuHmm` per GPU / scheduler thread
/// (`jtk_ctx` is thread-safe per distinct context, not re-entrant).
pub struct GpuHmm {
    ctx: *mut jtk_ctx,
}
unsafe impl Send for GpuHmm {}

fn offsets_u32<'a, I: Iterator<Item = &'a [u8]>>(items: I) -> (Vec<u8>, Vec<u32>) {
    let (mut cat, mut off) = (Vec::new(), vec![0u32]);
    for x in items {
        cat.extend_from_slice(x);
        off.push(cat.len() as u32);
    }
    (cat, off)
}
fn ops_bytes(ops: &[Op]) -> &[u8] {
    // Op is repr(u8) with the ABI's values
    unsafe { std::slice::from_raw_parts(ops.as_ptr() as *const u8, ops.len()) }
}

impl GpuHmm {
    /// `device < 0`: the current CUDA device.  Panics like the reference does on unrecoverable errors
    /// (`model_tune.rs:21`, `pseudo_mcmc.rs:100`).
    pub fn new(device: i32) -> GpuHmm {
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { jtk_ctx_create(device, 0, &mut ctx) };
        if rc != 0 {
            let msg = unsafe { CStr::from_ptr(jtk_last_error(std::ptr::null())) }.to_string_lossy().into_owned();
            panic!("jtk_ctx_create failed ({rc}): {msg}");
        }
        GpuHmm { ctx }
    }
    pub fn raw(&mut self) -> *mut jtk_ctx {
        self.ctx
    }
    fn check(&self, rc: c_int, what: &str) {
        if rc != 0 {
            let msg = unsafe { CStr::from_ptr(jtk_last_error(self.ctx)) }.to_string_lossy().into_owned();
            panic!("jtk_gpu {what} failed ({rc}): {msg}");
        }
    }

    /// K1. Drop-in for `hmm.modification_table_antidiagonal(template, read, ops, band) -> (Vec<f64>, f64)`
    /// (`pseudo_mcmc.rs:62-63`): absolute log-likelihoods, `table[j * NUM_ROW + row]`, and `lk`.
    pub fn modification_table_antidiagonal(
        &mut self, hmm: &HmmParams, template: &[u8], read: &[u8], ops: &[Op], band: usize,
    ) -> (Vec<f64>, f64) {
        let mut table = vec![0f64; (template.len() + 1) * NUM_ROW];
        let mut lk = 0f64;
        let toff = [0u32, template.len() as u32];
        let roff = [0u32, read.len() as u32];
        let ooff = [0u32, ops.len() as u32];
        let rc = unsafe {
            jtk_hmm_modtable_batch(
                self.ctx, hmm, hmm, 1, 1, template.as_ptr(), toff.as_ptr(), read.as_ptr(), roff.as_ptr(),
                ops_bytes(ops).as_ptr(), ooff.as_ptr(), [1u8].as_ptr(), [0u32].as_ptr(), band as c_int, &mut lk,
                table.as_mut_ptr(), [0u64].as_ptr(),
            )
        };
        self.check(rc, "modification_table_antidiagonal");
        (table, lk)
    }

    /// The per-chunk loop of `pseudo_mcmc::modification_table` (`pseudo_mcmc.rs:45-68`) in one call: every read of the
    /// pile-up against `template`, model chosen by `strands[i]` (`:58-61`), `lk` already subtracted (`:64`).
    pub fn modification_tables(
        &mut self, fwd: &HmmParams, rev: &HmmParams, template: &[u8], reads: &[&[u8]], ops: &[Vec<Op>],
        strands: &[bool], band: usize,
    ) -> Vec<Vec<f64>> {
        let n = reads.len();
        assert!(ops.len() == n && strands.len() == n);
        let (rcat, roff) = offsets_u32(reads.iter().copied());
        let (ocat, ooff) = offsets_u32(ops.iter().map(|o| ops_bytes(o)));
        let toff = [0u32, template.len() as u32];
        let strand: Vec<u8> = strands.iter().map(|&s| s as u8).collect();
        let tidx = vec![0u32; n];
        let per = (template.len() + 1) * NUM_ROW;
        let tab_off: Vec<u64> = (0..n).map(|i| (i * per) as u64).collect();
        let mut lk = vec![0f64; n];
        let mut flat = vec![0f64; n * per];
        let rc = unsafe {
            jtk_hmm_modtable_batch(
                self.ctx, fwd, rev, n as c_int, 1, template.as_ptr(), toff.as_ptr(), rcat.as_ptr(), roff.as_ptr(),
                ocat.as_ptr(), ooff.as_ptr(), strand.as_ptr(), tidx.as_ptr(), band as c_int, lk.as_mut_ptr(),
                flat.as_mut_ptr(), tab_off.as_ptr(),
            )
        };
        self.check(rc, "modification_tables");
        flat.chunks(per).zip(&lk).map(|(t, &l)| t.iter().map(|x| x - l).collect()).collect()
    }

    /// K2. Drop-in for `hmm.likelihood_antidiagonal_bootstrap(template, read, band) -> f64`
    /// (`likelihood_gains.rs:27-28,282-283,301-302`): no guide ops, the library derives the path.
    pub fn likelihood_antidiagonal_bootstrap(&mut self, hmm: &HmmParams, template: &[u8], read: &[u8], band: usize) -> f64 {
        self.likelihoods_antidiagonal_bootstrap(hmm, &[(template, read)], band)[0]
    }
    /// The same for many (template, read) pairs in one launch: the 200 calls of one calibration sample
    /// (`likelihood_gains.rs:275-306`) or all 180 000 of `estimate_gain` belong in one call.
    pub fn likelihoods_antidiagonal_bootstrap(&mut self, hmm: &HmmParams, pairs: &[(&[u8], &[u8])], band: usize) -> Vec<f64> {
        let n = pairs.len();
        let (tcat, toff) = offsets_u32(pairs.iter().map(|p| p.0));
        let (rcat, roff) = offsets_u32(pairs.iter().map(|p| p.1));
        let strand = vec![1u8; n];
        let tidx: Vec<u32> = (0..n as u32).collect();
        let mut lk = vec![0f64; n];
        let rc = unsafe {
            jtk_hmm_likelihood_batch(
                self.ctx, hmm, hmm, n as c_int, n as c_int, tcat.as_ptr(), toff.as_ptr(), rcat.as_ptr(), roff.as_ptr(),
                std::ptr::null(), std::ptr::null(), strand.as_ptr(), tidx.as_ptr(), band as c_int, lk.as_mut_ptr(),
            )
        };
        self.check(rc, "likelihood_antidiagonal_bootstrap");
        lk
    }

    /// K3. Drop-in for `models.polish_until_converge_antidiagonal(draft, seqs, &mut ops, strands, &config) -> Vec<u8>`
    /// (`local_clustering/mod.rs:106,155-156`, `model_tune.rs:143`, `consensus/mod.rs:477-483`).  `ops` are rewritten
    /// in place against the polished consensus, as kiley does.
    pub fn polish_until_converge_antidiagonal<T: AsRef<[u8]>>(
        &mut self, fwd: &HmmParams, rev: &HmmParams, draft: &[u8], seqs: &[T], ops: &mut [Vec<Op>], strands: &[bool],
        config: &PolishConfig,
    ) -> Vec<u8> {
        let mut drafts = [draft.to_vec()];
        let tidx = vec![0u32; seqs.len()];
        self.polish_chunks(fwd, rev, &mut drafts, seqs, ops, strands, &tidx, config);
        let [cons] = drafts;
        cons
    }
    /// The batched form the per-chunk driver should use: all pile-ups of a `local_clustering_selected` call at once
    /// (`local_clustering/mod.rs:64-72`); `drafts[c]` is replaced by its polished consensus, read `p` belongs to chunk
    /// `tmpl_idx[p]`.
    pub fn polish_chunks<T: AsRef<[u8]>>(
        &mut self, fwd: &HmmParams, rev: &HmmParams, drafts: &mut [Vec<u8>], seqs: &[T], ops: &mut [Vec<Op>],
        strands: &[bool], tmpl_idx: &[u32], config: &PolishConfig,
    ) -> Vec<i32> {
        let (n_chunks, n) = (drafts.len(), seqs.len());
        assert!(ops.len() == n && strands.len() == n && tmpl_idx.len() == n);
        let (dcat, doff) = offsets_u32(drafts.iter().map(|d| d.as_slice()));
        let (rcat, roff) = offsets_u32(seqs.iter().map(|s| s.as_ref()));
        // in/out ops: every pair owns a slot with room for the edits a polish can add (2x + 64 columns)
        let (mut ops_pos, mut ops_cap, mut n_ops) = (Vec::with_capacity(n), Vec::with_capacity(n), Vec::with_capacity(n));
        let mut total = 0u64;
        for o in ops.iter() {
            let cap = 2 * o.len() as u32 + 64;
            ops_pos.push(total);
            ops_cap.push(cap);
            n_ops.push(o.len() as u32);
            total += cap as u64;
        }
        let mut ops_buf = vec![0u8; total as usize];
        for (o, &p) in ops.iter().zip(&ops_pos) {
            ops_buf[p as usize..p as usize + o.len()].copy_from_slice(ops_bytes(o));
        }
        let (mut cons_pos, mut cons_cap) = (Vec::with_capacity(n_chunks), Vec::with_capacity(n_chunks));
        let mut ctotal = 0u64;
        for d in drafts.iter() {
            let cap = 2 * d.len() as u32 + 64;
            cons_pos.push(ctotal);
            cons_cap.push(cap);
            ctotal += cap as u64;
        }
        let mut out_cons = vec![0u8; ctotal as usize];
        let mut out_len = vec![0u32; n_chunks];
        let mut out_iters = vec![0i32; n_chunks];
        let strand: Vec<u8> = strands.iter().map(|&s| s as u8).collect();
        let rc = unsafe {
            jtk_polish_until_converge_batch(
                self.ctx, fwd, rev, n_chunks as c_int, dcat.as_ptr(), doff.as_ptr(), n as c_int, rcat.as_ptr(),
                roff.as_ptr(), ops_buf.as_mut_ptr(), ops_pos.as_ptr(), ops_cap.as_ptr(), n_ops.as_mut_ptr(),
                strand.as_ptr(), tmpl_idx.as_ptr(), config, out_cons.as_mut_ptr(), cons_pos.as_ptr(), cons_cap.as_ptr(),
                out_len.as_mut_ptr(), out_iters.as_mut_ptr(),
            )
        };
        self.check(rc, "polish_until_converge_antidiagonal");
        for (p, o) in ops.iter_mut().enumerate() {
            let s = &ops_buf[ops_pos[p] as usize..ops_pos[p] as usize + n_ops[p] as usize];
            o.clear();
            o.extend(s.iter().map(|&b| match b {
                0 => Op::Match,
                1 => Op::Mismatch,
                2 => Op::Ins,
                _ => Op::Del,
            }));
        }
        for (c, d) in drafts.iter_mut().enumerate() {
            let s = &out_cons[cons_pos[c] as usize..cons_pos[c] as usize + out_len[c] as usize];
            d.clear();
            d.extend_from_slice(s);
        }
        out_iters
    }

    /// K4. Drop-in for `models.fit_antidiagonal_par_multiple(&packs, radius)` (`model_tune.rs:151`): one Baum-Welch
    /// update of both strand models in place.
    pub fn fit_antidiagonal_par_multiple<T: AsRef<[u8]>>(
        &mut self, fwd: &mut HmmParams, rev: &mut HmmParams, packs: &[TrainingDataPack<'_, T>], radius: usize,
    ) {
        let (tcat, toff) = offsets_u32(packs.iter().map(|p| p.consensus));
        let (rcat, roff) = offsets_u32(packs.iter().flat_map(|p| p.sequences.iter().map(|s| s.as_ref())));
        let (ocat, ooff) = offsets_u32(packs.iter().flat_map(|p| p.operations.iter().map(|o| ops_bytes(o))));
        let strand: Vec<u8> = packs.iter().flat_map(|p| p.directions.iter().map(|&s| s as u8)).collect();
        let tidx: Vec<u32> =
            packs.iter().enumerate().flat_map(|(t, p)| std::iter::repeat(t as u32).take(p.sequences.len())).collect();
        let rc = unsafe {
            jtk_hmm_fit_batch(
                self.ctx, fwd, rev, tidx.len() as c_int, packs.len() as c_int, tcat.as_ptr(), toff.as_ptr(), rcat.as_ptr(),
                roff.as_ptr(), ocat.as_ptr(), ooff.as_ptr(), strand.as_ptr(), tidx.as_ptr(), radius as c_int,
            )
        };
        self.check(rc, "fit_antidiagonal_par_multiple");
    }
}

impl Drop for GpuHmm {
    fn drop(&mut self) {
        unsafe { jtk_ctx_destroy(self.ctx) }
    }
}
