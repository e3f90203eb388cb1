"""`PoseidonTranscript<G1Affine, NativeLoader, ..>` -- oracle restatement.  TEST INFRASTRUCTURE ONLY.

Follows snark-verifier/src/system/halo2/transcript/halo2.rs:163-304 (native half) and the
`Transcript*` traits of snark-verifier/src/util/transcript.rs:9-62.
"""
from . import bn254
from .bn254 import R
from .loader import EcPoint, Scalar, fe_to_fe
from .poseidon import Poseidon


class VerifyError(Exception):
    """snark-verifier/src/lib.rs:21-30 `Error`; `kind` is one of
    InvalidInstances | InvalidProtocol | AssertionFailure | Transcript"""

    def __init__(self, kind, msg=""):
        super().__init__(f"{kind}: {msg}")
        self.kind = kind
        self.msg = msg


class ReferencePanic(Exception):
    """The reference does not return an `Err` here: it panics (`unwrap()` / `assert!`)."""


class PoseidonTranscript:
    def __init__(self, loader, stream=b"", spec=None):
        self.loader = loader
        self.stream = bytes(stream)
        self.pos = 0
        self.buf = Poseidon(spec)
        self.out = bytearray()  # TranscriptWrite side

    def new_stream(self, stream):
        """halo2.rs:185-189"""
        self.buf.clear()
        self.stream = bytes(stream)
        self.pos = 0
        tr = self.loader.tracer
        if tr is not None:
            tr.emit("t_clear")

    # ---- Transcript (halo2.rs:198-227)
    def squeeze_challenge(self):
        v = self.buf.squeeze()
        tr = self.loader.tracer
        reg = None
        if tr is not None:
            reg = tr.new_s()
            tr.emit("t_squeeze", reg)
        return Scalar(v, self.loader, reg)

    def squeeze_n_challenges(self, n):
        return [self.squeeze_challenge() for _ in range(n)]

    def common_scalar(self, scalar):
        self.buf.update([scalar.v])
        tr = self.loader.tracer
        if tr is not None:
            tr.emit("t_common_scalar", scalar.reg)

    def common_ec_point(self, ec_point):
        if ec_point.pt is None:  # `coordinates()` is None for the identity (halo2.rs:215-224)
            raise VerifyError("Transcript", "Invalid elliptic curve point encoding in proof")
        x, y = ec_point.pt
        self.buf.update([fe_to_fe(x), fe_to_fe(y)])
        tr = self.loader.tracer
        if tr is not None:
            tr.emit("t_common_point", ec_point.reg)

    # ---- TranscriptRead (halo2.rs:229-261)
    def _read(self, n):
        if self.pos + n > len(self.stream):
            raise VerifyError("Transcript", "failed to fill whole buffer")  # io::ErrorKind::UnexpectedEof
        d = self.stream[self.pos : self.pos + n]
        self.pos += n
        return d

    def read_scalar(self):
        data = self._read(32)
        v = bn254.fr_from_bytes(data)
        if v is None:
            raise VerifyError("Transcript", "Invalid scalar encoding in proof")
        tr = self.loader.tracer
        reg = None
        if tr is not None:
            reg = tr.new_s()
            tr.emit("t_read_scalar", reg)
        s = Scalar(v, self.loader, reg)
        self.buf.update([v])
        return s

    def read_n_scalars(self, n):
        return [self.read_scalar() for _ in range(n)]

    def read_ec_point(self):
        data = self._read(32)
        ok, pt = bn254.g1_from_bytes(data)
        if not ok:
            raise VerifyError("Transcript", "Invalid elliptic curve point encoding in proof")
        tr = self.loader.tracer
        reg = None
        if tr is not None:
            reg = tr.new_p()
            tr.emit("t_read_point", reg)
        p = EcPoint(pt, self.loader, reg)
        if pt is None:
            raise VerifyError("Transcript", "Invalid elliptic curve point encoding in proof")
        x, y = pt
        self.buf.update([fe_to_fe(x), fe_to_fe(y)])
        return p

    def read_n_ec_points(self, n):
        return [self.read_ec_point() for _ in range(n)]

    # ---- TranscriptWrite (halo2.rs:279-304)
    def write_scalar(self, v):
        self.buf.update([v % R])
        self.out += bn254.fe_to_bytes(v % R)

    def write_ec_point(self, pt):
        if pt is None:
            raise VerifyError("Transcript", "Invalid elliptic curve point encoding in proof")
        self.buf.update([fe_to_fe(pt[0]), fe_to_fe(pt[1])])
        self.out += bn254.g1_to_bytes(pt)

    def finalize(self):
        return bytes(self.out)


class EvmTranscript:
    """`EvmTranscript<G1Affine, NativeLoader, S, Vec<u8>>` (snark-verifier/src/system/halo2/transcript/evm.rs:152-243):
    Keccak-256 over a byte buffer; scalars 32 B big-endian, points uncompressed x || y 64 B big-endian."""

    def __init__(self, loader, stream=b""):
        from .keccak import keccak256

        self._keccak = keccak256
        self.loader = loader
        self.stream = bytes(stream)
        self.pos = 0
        self.buf = b""
        self.out = bytearray()

    def squeeze_challenge(self):
        """evm.rs:172-182"""
        data = self.buf + (b"\x01" if len(self.buf) == 0x20 else b"")
        h = self._keccak(data)
        self.buf = h
        return Scalar(int.from_bytes(h, "big") % R, self.loader)

    def squeeze_n_challenges(self, n):
        return [self.squeeze_challenge() for _ in range(n)]

    def common_scalar(self, scalar):
        self.buf += int(scalar.v).to_bytes(32, "big")

    def common_ec_point(self, ec_point):
        if ec_point.pt is None:
            raise VerifyError("Transcript", "Invalid elliptic curve point")
        self.buf += int(ec_point.pt[0]).to_bytes(32, "big") + int(ec_point.pt[1]).to_bytes(32, "big")

    def _read(self, n):
        if self.pos + n > len(self.stream):
            raise VerifyError("Transcript", "failed to fill whole buffer")
        d = self.stream[self.pos : self.pos + n]
        self.pos += n
        return d

    def read_scalar(self):
        v = int.from_bytes(self._read(32), "big")
        if v >= R:
            raise VerifyError("Transcript", "Invalid scalar encoding in proof")
        s = Scalar(v, self.loader)
        self.common_scalar(s)
        return s

    def read_n_scalars(self, n):
        return [self.read_scalar() for _ in range(n)]

    def read_ec_point(self):
        x = int.from_bytes(self._read(32), "big")
        y = int.from_bytes(self._read(32), "big")
        if x >= bn254.P or y >= bn254.P:
            raise VerifyError("Transcript", "Invalid elliptic curve point encoding in proof")
        pt = None if (x == 0 and y == 0) else (x, y)  # halo2curves `from_xy` accepts (0,0) as the identity ...
        if pt is not None and not bn254.g1_is_on_curve(pt):
            raise VerifyError("Transcript", "Invalid elliptic curve point encoding in proof")
        p = EcPoint(pt, self.loader)
        self.common_ec_point(p)  # ... which `common_ec_point` then rejects
        return p

    def read_n_ec_points(self, n):
        return [self.read_ec_point() for _ in range(n)]

    # TranscriptWrite (evm.rs:262-296)
    def write_scalar(self, v):
        self.buf += int(v % R).to_bytes(32, "big")
        self.out += int(v % R).to_bytes(32, "big")

    def write_ec_point(self, pt):
        if pt is None:
            raise VerifyError("Transcript", "Invalid elliptic curve point")
        b = int(pt[0]).to_bytes(32, "big") + int(pt[1]).to_bytes(32, "big")
        self.buf += b
        self.out += b

    def finalize(self):
        return bytes(self.out)


def make_transcript(kind, loader, stream=b""):
    return PoseidonTranscript(loader, stream) if kind == "poseidon" else EvmTranscript(loader, stream)
