//! `snark-verifier-cuda`: the NativeLoader KZG/PLONK verification path of `snark-verifier` on a B200 through libsvk.
//!
//! * [`CudaPlonkVerifier`]: batch API (`succinct_verify_batch`, `verify_batch`, `kzg_as_create_proof`, `kzg_as_verify_zk`, `decide_all`).
//! * `impl SnarkVerifier<G1Affine, NativeLoader> for CudaPlonkVerifier<MOS>`: the reference's single-proof interface
//!   (`snark-verifier/src/verifier.rs:13-44`); `Proof` is the raw transcript bytes, parsing happens on the device.
//!
//! The statuses libsvk reports per proof map one-to-one onto `snark_verifier::Error` (`src/lib.rs:21-30`).
mod ffi;

use std::{
    collections::HashMap,
    ffi::CStr,
    marker::PhantomData,
    ptr,
    sync::{Arc, Mutex},
};

use itertools::Itertools;
use snark_verifier::{
    halo2_curves::{
        bn256::{Bn256, Fq, Fq2, Fr, G1Affine, G2Affine},
        ff::PrimeField,
        group::GroupEncoding,
        CurveAffine,
    },
    loader::native::NativeLoader,
    pcs::kzg::{Bdfg21, Gwc19, KzgAccumulator, KzgDecidingKey},
    util::transcript::TranscriptRead,
    verifier::{plonk::PlonkProtocol, SnarkVerifier},
    Error,
};

/// Multi-open scheme selector (`pcs/kzg/multiopen/{bdfg21,gwc19}.rs`).
pub trait Mos: 'static {
    const ID: i32;
}
impl Mos for Bdfg21 {
    const ID: i32 = 0;
}
impl Mos for Gwc19 {
    const ID: i32 = 1;
}

/// The triple `KzgDecidingKey::new` takes (`pcs/kzg/decider.rs:17-25`); its fields are private in the reference, so the
/// binding keeps its own copy.
#[derive(Clone, Copy, Debug)]
pub struct CudaDecidingKey {
    pub g1: G1Affine,
    pub g2: G2Affine,
    pub s_g2: G2Affine,
}

impl From<(G1Affine, G2Affine, G2Affine)> for CudaDecidingKey {
    fn from((g1, g2, s_g2): (G1Affine, G2Affine, G2Affine)) -> Self {
        Self { g1, g2, s_g2 }
    }
}

impl CudaDecidingKey {
    pub fn to_reference(&self) -> KzgDecidingKey<Bn256> {
        KzgDecidingKey::new(self.g1, self.g2, self.s_g2)
    }
}

fn fe<F: PrimeField<Repr = [u8; 32]>>(x: &F) -> ffi::svk_fe {
    ffi::svk_fe { b: x.to_repr() }
}

fn g1(p: &G1Affine) -> ffi::svk_g1 {
    // (0, 0) encodes the identity at the ABI
    match Option::<_>::from(p.coordinates()) {
        Some(c) => {
            let c: snark_verifier::halo2_curves::Coordinates<G1Affine> = c;
            ffi::svk_g1 { x: fe::<Fq>(c.x()), y: fe::<Fq>(c.y()) }
        }
        None => ffi::svk_g1::default(),
    }
}

fn fq2(v: &Fq2) -> (ffi::svk_fe, ffi::svk_fe) {
    (fe(&v.c0), fe(&v.c1))
}

fn g2(p: &G2Affine) -> ffi::svk_g2 {
    let c = p.coordinates().unwrap();
    let (x0, x1) = fq2(c.x());
    let (y0, y1) = fq2(c.y());
    ffi::svk_g2 { x_c0: x0, x_c1: x1, y_c0: y0, y_c1: y1 }
}

fn g1_from(p: &ffi::svk_g1) -> G1Affine {
    if p.x.b == [0u8; 32] && p.y.b == [0u8; 32] {
        return G1Affine::identity();
    }
    let x = Fq::from_repr(p.x.b).unwrap();
    let y = Fq::from_repr(p.y.b).unwrap();
    G1Affine::from_xy(x, y).unwrap()
}

/// Status word -> the reference's error (`snark-verifier/src/lib.rs:21-30`).
pub fn status_to_result(status: i32) -> Result<(), Error> {
    let sub = status >> 8;
    match status & 0xff {
        ffi::SVK_OK => Ok(()),
        ffi::SVK_INVALID_INSTANCES => Err(Error::InvalidInstances),
        ffi::SVK_INVALID_PROTOCOL => Err(Error::InvalidProtocol("protocol rejected by the verifier".to_string())),
        ffi::SVK_ASSERTION_FAILURE => Err(Error::AssertionFailure("e(lhs, g2)·e(rhs, -s_g2) == O".to_string())),
        ffi::SVK_TRANSCRIPT => Err(match sub {
            ffi::SVK_T_EOF => Error::Transcript(std::io::ErrorKind::UnexpectedEof, "failed to fill whole buffer".to_string()),
            ffi::SVK_T_SCALAR_RANGE => {
                Error::Transcript(std::io::ErrorKind::Other, "Invalid scalar encoding in proof".to_string())
            }
            ffi::SVK_T_POINT_INVALID => {
                Error::Transcript(std::io::ErrorKind::Other, "Invalid elliptic curve point encoding in proof".to_string())
            }
            _ => Error::Transcript(std::io::ErrorKind::Other, "Cannot write points at infinity to the transcript".to_string()),
        }),
        // the reference panics (`from_xy(..).unwrap()`, pcs/kzg/accumulator.rs:72-73); a service would rather get an error
        _ => Err(Error::AssertionFailure("old accumulator limbs do not decode to a curve point".to_string())),
    }
}

struct Ctx(*mut ffi::svk_ctx);
// libsvk serialises calls on one context with an internal mutex (include/svk.h "Threading")
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {}
impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { ffi::svk_destroy(self.0) }
    }
}

fn last_error(ctx: *mut ffi::svk_ctx) -> String {
    unsafe { CStr::from_ptr(ffi::svk_last_error(ctx)).to_string_lossy().into_owned() }
}

/// One compiled (deciding key, protocol, multi-open scheme) on one GPU.
pub struct CudaPlonkVerifier<MOS> {
    ctx: Arc<Ctx>,
    dk_id: i32,
    proto_id: i32,
    n_instances: usize,
    n_old: usize,
    num_instance: Vec<usize>,
    _m: PhantomData<MOS>,
}

/// A snark of the protocol the verifier was compiled for.
pub struct SnarkRef<'a> {
    pub instances: &'a [Vec<Fr>],
    pub proof: &'a [u8],
}

impl<MOS: Mos> CudaPlonkVerifier<MOS> {
    /// `dk` as built at `examples/recursion.rs:847`; `protocol` as returned by `compile()` (`system/halo2.rs:82`) or read
    /// from a `Snark` file.  `evm_transcript`: the proofs were written with the Keccak `EvmTranscript`.
    pub fn new(device: i32, dk: &CudaDecidingKey, protocol: &PlonkProtocol<G1Affine>, evm_transcript: bool) -> Result<Self, String> {
        let mut raw: *mut ffi::svk_ctx = ptr::null_mut();
        if unsafe { ffi::svk_create(device, &mut raw) } != 0 {
            return Err(last_error(ptr::null_mut()));
        }
        let ctx = Arc::new(Ctx(raw));
        let key = ffi::svk_deciding_key { g1: g1(&dk.g1), g2: g2(&dk.g2), s_g2: g2(&dk.s_g2) };
        let dk_id = unsafe { ffi::svk_dk_load(ctx.0, &key) };
        if dk_id < 0 {
            return Err(last_error(ctx.0));
        }
        // `PlonkProtocol` derives Serialize (verifier/plonk/protocol.rs:19): the reference's own wire form goes in as it is
        let bytes = bincode::serialize(protocol).map_err(|e| e.to_string())?;
        let proto_id = unsafe {
            ffi::svk_protocol_compile_bincode(ctx.0, bytes.as_ptr(), bytes.len(), 0, MOS::ID, evm_transcript as i32, dk_id, ptr::null_mut(), ptr::null_mut())
        };
        if proto_id < 0 {
            return Err(last_error(ctx.0));
        }
        let mut info = [0u32; 20];
        unsafe { ffi::svk_protocol_info(ctx.0, proto_id, info.as_mut_ptr()) };
        Ok(Self {
            ctx,
            dk_id,
            proto_id,
            n_instances: info[1] as usize,
            n_old: info[16] as usize,
            num_instance: protocol.num_instance.clone(),
            _m: PhantomData,
        })
    }

    /// Row-major packing for the ABI.  Returns the indices of snarks whose instance columns do not match
    /// `protocol.num_instance` (`verifier/plonk/proof.rs:66-69`): they are reported as `InvalidInstances`.
    fn pack(&self, snarks: &[SnarkRef<'_>]) -> (Vec<ffi::svk_fe>, Vec<u8>, Vec<u32>, usize, Vec<usize>) {
        let stride = snarks.iter().map(|s| s.proof.len()).max().unwrap_or(0).max(32);
        let stride = (stride + 31) / 32 * 32;
        let mut inst = vec![ffi::svk_fe::default(); snarks.len() * self.n_instances.max(1)];
        let mut proofs = vec![0u8; snarks.len() * stride];
        let mut lens = Vec::with_capacity(snarks.len());
        let mut bad = Vec::new();
        for (i, s) in snarks.iter().enumerate() {
            if s.instances.iter().map(|c| c.len()).collect_vec() != self.num_instance {
                bad.push(i);
            } else {
                for (k, x) in s.instances.iter().flatten().enumerate() {
                    inst[i * self.n_instances + k] = fe(x);
                }
            }
            proofs[i * stride..i * stride + s.proof.len()].copy_from_slice(s.proof);
            lens.push(s.proof.len() as u32);
        }
        (inst, proofs, lens, stride, bad)
    }

    /// Batch twin of `PlonkSuccinctVerifier::verify` (`verifier/plonk.rs:58-93`): per snark the accumulators
    /// `[new, old_0, ..]` or the error `read_proof` / `verify` would have returned.
    pub fn succinct_verify_batch(&self, snarks: &[SnarkRef<'_>]) -> Vec<Result<Vec<KzgAccumulator<G1Affine, NativeLoader>>, Error>> {
        let n = snarks.len();
        if n == 0 {
            return Vec::new();
        }
        let (inst, proofs, lens, stride, bad) = self.pack(snarks);
        let apk = 1 + self.n_old;
        let mut accs = vec![ffi::svk_acc::default(); n * apk];
        let mut status = vec![0i32; n];
        let rc = unsafe {
            ffi::svk_plonk_succinct_verify_batch(
                self.ctx.0, self.proto_id, n, inst.as_ptr(), self.n_instances as u32, proofs.as_ptr(), stride, lens.as_ptr(),
                accs.as_mut_ptr(), ptr::null_mut(), status.as_mut_ptr(),
            )
        };
        assert_eq!(rc, 0, "libsvk: {}", last_error(self.ctx.0));
        for i in bad {
            status[i] = ffi::SVK_INVALID_INSTANCES;
        }
        (0..n)
            .map(|i| {
                status_to_result(status[i]).map(|_| {
                    accs[i * apk..(i + 1) * apk].iter().map(|a| KzgAccumulator::new(g1_from(&a.lhs), g1_from(&a.rhs))).collect_vec()
                })
            })
            .collect()
    }

    /// Batch twin of `PlonkVerifier::verify` (`verifier/plonk.rs:125-134`): succinct-verify each snark, fold the
    /// accumulators with `KzgAs` (`pcs/kzg/accumulation.rs:139-196`, zk = false; groups of `group_size`, 0 = the flat fold of
    /// `sdk/src/halo2/aggregation.rs:235-245`), decide ONCE (`pcs/kzg/decider.rs:60-68`).  When the single pairing
    /// rejects, every accumulator is decided to name the culprits, as per-proof `decide_all` would.
    pub fn verify_batch(&self, snarks: &[SnarkRef<'_>], group_size: usize) -> Vec<Result<(), Error>> {
        let n = snarks.len();
        if n == 0 {
            return Vec::new();
        }
        let (inst, proofs, lens, stride, bad) = self.pack(snarks);
        let mut status = vec![0i32; n];
        let mut folded = ffi::svk_acc::default();
        let mut ok = 0u8;
        let rc = unsafe {
            ffi::svk_plonk_verify_batch(
                self.ctx.0, self.proto_id, n, inst.as_ptr(), self.n_instances as u32, proofs.as_ptr(), stride, lens.as_ptr(),
                group_size, 1, status.as_mut_ptr(), &mut folded, &mut ok,
            )
        };
        assert_eq!(rc, 0, "libsvk: {}", last_error(self.ctx.0));
        for i in bad {
            status[i] = ffi::SVK_INVALID_INSTANCES;
        }
        status.into_iter().map(status_to_result).collect()
    }

    /// `KzgAs::create_proof` / `verify` with zk = false (`pcs/kzg/accumulation.rs:29-62, 139-196`): the folded accumulator
    /// and the root challenge r.
    pub fn kzg_as_create_proof(&self, accs: &[KzgAccumulator<G1Affine, NativeLoader>], group_size: usize) -> Result<(KzgAccumulator<G1Affine, NativeLoader>, Fr), Error> {
        let packed = accs.iter().map(|a| ffi::svk_acc { lhs: g1(&a.lhs), rhs: g1(&a.rhs) }).collect_vec();
        let (mut out, mut r, mut status) = (ffi::svk_acc::default(), ffi::svk_fe::default(), 0i32);
        let rc = unsafe { ffi::svk_kzg_as_fold(self.ctx.0, packed.len(), packed.as_ptr(), group_size, &mut out, &mut r, &mut status) };
        assert_eq!(rc, 0, "libsvk: {}", last_error(self.ctx.0));
        status_to_result(status)?;
        Ok((KzgAccumulator::new(g1_from(&out.lhs), g1_from(&out.rhs)), Fr::from_repr(r.b).unwrap()))
    }

    /// `KzgAs::{read_proof, verify}` with `KzgAsVerifyingKey(true)` (zk accumulation, `pcs/kzg/accumulation.rs:29-62, 124-133`):
    /// `as_proof` = the two compressed blind points a zk `create_proof` wrote; they are absorbed after the instances and folded
    /// in with r^n.  A short proof, a bad or an identity point come back as `Error::Transcript`, like `read_ec_point`.
    pub fn kzg_as_verify_zk(&self, accs: &[KzgAccumulator<G1Affine, NativeLoader>], as_proof: &[u8]) -> Result<(KzgAccumulator<G1Affine, NativeLoader>, Fr), Error> {
        let packed = accs.iter().map(|a| ffi::svk_acc { lhs: g1(&a.lhs), rhs: g1(&a.rhs) }).collect_vec();
        let (mut out, mut r, mut status) = (ffi::svk_acc::default(), ffi::svk_fe::default(), 0i32);
        let rc = unsafe {
            ffi::svk_kzg_as_fold_zk(self.ctx.0, packed.len(), packed.as_ptr(), as_proof.as_ptr(), as_proof.len(), &mut out, &mut r, &mut status)
        };
        assert_eq!(rc, 0, "libsvk: {}", last_error(self.ctx.0));
        status_to_result(status)?;
        Ok((KzgAccumulator::new(g1_from(&out.lhs), g1_from(&out.rhs)), Fr::from_repr(r.b).unwrap()))
    }

    /// `KzgAs::decide_all` (`pcs/kzg/decider.rs:70-81`).
    pub fn decide_all(&self, accs: &[KzgAccumulator<G1Affine, NativeLoader>]) -> Result<(), Error> {
        let packed = accs.iter().map(|a| ffi::svk_acc { lhs: g1(&a.lhs), rhs: g1(&a.rhs) }).collect_vec();
        let mut ok = vec![0u8; packed.len()];
        let rc = unsafe { ffi::svk_kzg_decide_batch(self.ctx.0, self.dk_id, packed.len(), packed.as_ptr(), ok.as_mut_ptr()) };
        assert_eq!(rc, 0, "libsvk: {}", last_error(self.ctx.0));
        if ok.iter().all(|&b| b == 1) {
            Ok(())
        } else {
            status_to_result(ffi::SVK_ASSERTION_FAILURE)
        }
    }
}

/// Compiled verifiers of the single-proof trait interface, keyed by (scheme, bincode of key and protocol).
fn cache() -> &'static Mutex<HashMap<(i32, Vec<u8>), Arc<dyn std::any::Any + Send + Sync>>> {
    static CACHE: std::sync::OnceLock<Mutex<HashMap<(i32, Vec<u8>), Arc<dyn std::any::Any + Send + Sync>>>> = std::sync::OnceLock::new();
    CACHE.get_or_init(|| Mutex::new(HashMap::new()))
}

unsafe impl<MOS> Send for CudaPlonkVerifier<MOS> {}
unsafe impl<MOS> Sync for CudaPlonkVerifier<MOS> {}

/// Single-proof drop-in with the reference's interface: call sites written against
/// `PlonkVerifier::<KzgAs<Bn256, MOS>>::{read_proof, verify}` keep compiling with `CudaPlonkVerifier::<MOS>`.
impl<MOS: Mos> SnarkVerifier<G1Affine, NativeLoader> for CudaPlonkVerifier<MOS> {
    type VerifyingKey = CudaDecidingKey;
    type Protocol = PlonkProtocol<G1Affine>;
    /// The transcript bytes in reading order; they are parsed (and checked) on the device.
    type Proof = Vec<u8>;
    type Output = ();

    /// Reads exactly what `PlonkProof::read` reads (`verifier/plonk/proof.rs:52-153`) and re-encodes it: compressed points
    /// (`G1Affine::to_bytes`) and `to_repr()` scalars, the layout the native `PoseidonTranscript` reads
    /// (`system/halo2/transcript/halo2.rs:232-261`).
    fn read_proof<T>(_: &Self::VerifyingKey, protocol: &Self::Protocol, instances: &[Vec<Fr>], transcript: &mut T) -> Result<Vec<u8>, Error>
    where
        T: TranscriptRead<G1Affine, NativeLoader>,
    {
        if let Some(state) = &protocol.transcript_initial_state {
            transcript.common_scalar(state)?;
        }
        if protocol.num_instance != instances.iter().map(|i| i.len()).collect_vec() {
            return Err(Error::InvalidInstances);
        }
        for x in instances.iter().flatten() {
            transcript.common_scalar(x)?;
        }
        let mut bytes = Vec::new();
        let mut points = |t: &mut T, n: usize, out: &mut Vec<u8>| -> Result<(), Error> {
            for p in t.read_n_ec_points(n)? {
                out.extend_from_slice(p.to_bytes().as_ref());
            }
            Ok(())
        };
        for (&n, &m) in protocol.num_witness.iter().zip(protocol.num_challenge.iter()) {
            points(transcript, n, &mut bytes)?;
            transcript.squeeze_n_challenges(m);
        }
        points(transcript, protocol.quotient.num_chunk(), &mut bytes)?;
        transcript.squeeze_challenge(); // z
        for e in transcript.read_n_scalars(protocol.evaluations.len())? {
            bytes.extend_from_slice(e.to_repr().as_ref());
        }
        if MOS::ID == 0 {
            // Bdfg21Proof::read (pcs/kzg/multiopen/bdfg21.rs:101-114)
            transcript.squeeze_n_challenges(2); // mu, gamma
            points(transcript, 1, &mut bytes)?; // w
            transcript.squeeze_challenge(); // z'
            points(transcript, 1, &mut bytes)?; // w'
        } else {
            // Gwc19Proof::read (pcs/kzg/multiopen/gwc19.rs:100-108): one opening proof per distinct rotation
            transcript.squeeze_challenge(); // v
            let sets = protocol.queries.iter().map(|q| q.rotation).unique().count();
            points(transcript, sets, &mut bytes)?;
            transcript.squeeze_challenge(); // u
        }
        Ok(bytes)
    }

    fn verify(vk: &Self::VerifyingKey, protocol: &Self::Protocol, instances: &[Vec<Fr>], proof: &Vec<u8>) -> Result<(), Error> {
        let mut key = bincode::serialize(protocol).map_err(|e| Error::InvalidProtocol(e.to_string()))?;
        key.extend_from_slice(vk.g1.to_bytes().as_ref());
        key.extend_from_slice(vk.g2.to_bytes().as_ref());
        key.extend_from_slice(vk.s_g2.to_bytes().as_ref());
        let v: Arc<CudaPlonkVerifier<MOS>> = {
            let mut c = cache().lock().unwrap();
            let e = match c.get(&(MOS::ID, key.clone())) {
                Some(e) => e.clone(),
                None => {
                    let made: Arc<dyn std::any::Any + Send + Sync> =
                        Arc::new(CudaPlonkVerifier::<MOS>::new(0, vk, protocol, false).map_err(Error::InvalidProtocol)?);
                    c.insert((MOS::ID, key), made.clone());
                    made
                }
            };
            e.downcast::<CudaPlonkVerifier<MOS>>().expect("cache entry of another scheme")
        };
        v.verify_batch(&[SnarkRef { instances, proof }], 0).pop().unwrap()
    }
}
