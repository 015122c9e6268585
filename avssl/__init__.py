"""Drop-in surface of atosystem/SpeechCLIP's ``avssl`` package for its ONE hot path: the same module
paths, class names, constructor arguments and return conventions as the reference, with the arithmetic running as
sm_100a CUDA kernels behind ``libspeechclip_b200.so`` (speechclip_b200/).  Data loading, task runners and logging are
out of scope."""
