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
