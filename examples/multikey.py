#!/usr/bin/env python
"""The workload of the reference's examples/multikey.jl on the B200 engine: MK-TFHE NAND between ciphertexts encrypted
under the keys of several parties, ten random trials (plus the whole batch of trials in ONE library call).

    python examples/multikey.py [parties]       # 2 (default), 4 or 8; needs a B200
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tfhe_jl_b200 as T  # noqa: E402


def main(parties=2, trials=10, seed=None):
    params = {2: T.mktfhe_parameters_2party, 4: T.mktfhe_parameters_4party, 8: T.mktfhe_parameters_8party}[parties]
    rng = np.random.default_rng(seed)
    secret_keys = [T.SecretKey(rng, params) for _ in range(parties)]             # on the clients' machines
    shared_key = T.SharedKey(rng, params)                                        # created by the server
    ck_parts = [T.CloudKeyPart(rng, sk, shared_key) for sk in secret_keys]       # on the clients' machines
    cloud_key = T.MKCloudKey(ck_parts)                                           # on the server: expansion + transform on the GPU

    ok = 0
    mess = rng.integers(0, 2, (trials, 2)).astype(bool)
    for trial in range(trials):
        mess1, mess2 = bool(mess[trial, 0]), bool(mess[trial, 1])
        enc1, enc2 = T.mk_encrypt(rng, secret_keys, mess1), T.mk_encrypt(rng, secret_keys, mess2)
        assert T.mk_decrypt(secret_keys, enc1) == mess1 and T.mk_decrypt(secret_keys, enc2) == mess2
        dec_out = T.mk_decrypt(secret_keys, T.mk_gate_nand(cloud_key, enc1, enc2))
        ok += dec_out == (not (mess1 and mess2))
        print(f"Trial {trial + 1}: {mess1} NAND {mess2} = {dec_out}")
    # the batched form: all trials in one call
    out = T.mk_gate_nand(cloud_key, T.mk_encrypt(rng, secret_keys, mess[:, 0]), T.mk_encrypt(rng, secret_keys, mess[:, 1]))
    batch_ok = int((T.mk_decrypt(secret_keys, out) == ~(mess[:, 0] & mess[:, 1])).sum())
    print(f"{ok}/{trials} single gates and {batch_ok}/{trials} gates of the batched call decrypt to the plaintext NAND")
    return ok, batch_ok


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 2)
