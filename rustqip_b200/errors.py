"""Error type of the host-side mirror.

Mirrors ``CircuitError`` / ``CircuitResult`` of the reference
(qip/src/errors.rs:6-22): one generic error carrying a message.  The C ABI
reports failures as status codes; the mirror turns every non-zero status into
this exception with the library's ``qipb200_last_error`` text.
"""


class CircuitError(Exception):
    """qip::errors::CircuitError::Generic(String) (qip/src/errors.rs:6-22)."""

    def __init__(self, msg, status=None):
        super().__init__(msg)
        self.msg = msg
        self.status = status


class B200Unavailable(RuntimeError):
    """Raised when libqipb200.so or a CUDA device is missing.  There is no CPU path."""
