#!/usr/bin/env python3
"""Generate the committed golden fixtures under tests/golden/.

Ground truth is Python's arbitrary-precision pow()/% -- every function on the
hot path returns a canonical residue, so this is what the reference's
ippModExp / encrypt / decrypt / + / * return on the same inputs.  The ISO/IEC
18033-6 constants are the ones the reference's own known-answer test holds
(/root/reference/test/test_cryptography.cpp:104-196); this script re-derives
c1, c2, c1c2, m1m2 from p, q, r0, r1 and the two plaintexts and asserts they
equal the published values before writing anything, so the fixture is pinned
to the reference's vectors and not merely to itself.

Run:  python tests/golden/make_golden.py       (deterministic, ~1 min)
"""
import json
import math
import os
import random

HERE = os.path.dirname(os.path.abspath(__file__))

# --- ISO/IEC 18033-6 vectors, test/test_cryptography.cpp:104-196 -------------
ISO = dict(
    p=int("ff03b1a74827c746db83d2eaff00067622f545b62584321256e62b01509f10962f9c5c"
          "8fd0b7f5184a9ce8e81f439df47dda14563dd55a221799d2aa57ed2713271678a5a0b8b4"
          "0a84ad13d5b6e6599e6467c670109cf1f45ccfed8f75ea3b814548ab294626fe4d14ff76"
          "4dd8b091f11a0943a2dd2b983b0df02f4c4d00b413", 16),
    q=int("dacaabc1dc57faa9fd6a4274c4d588765a1d3311c22e57d8101431b07eb3ddcb05d77d"
          "9a742ac2322fe6a063bd1e05acb13b0fe91c70115c2b1eee1155e072527011a5f849de70"
          "72a1ce8e6b71db525fbcda7a89aaed46d27aca5eaeaf35a26270a4a833c5cda681ffd49b"
          "aa0f610bad100cdf47cc86e5034e2a0b2179e04ec7", 16),
    c1=int("1fb7f08a42deb47876e4cbdc3f0b172c033563a696ad7a7c76fa5971b793fa488dcdd6"
           "bd65c7c5440d67d847cb89ccca468b2c96763fff5a5ece8330251112d65e59b7da94cfe9"
           "309f441ccc8f59c67dec75113d37b1ee929c8d4ce6b5e561a30a91104b0526de892e4eff"
           "9f4fbecba3db8ed94267be31df360feaffb1151ef5b5a8e51777f09d38072bcb1b1ad15d"
           "80d5448fd0edb41cc499f8eebae2af26569427a26d0afeaa833173d6ae4e5f84eb88c0c6"
           "8c29baecf7ec5af2c1c5577336ca9482690f1c94597654afda84c6fb74df95cdd08fa9a6"
           "6296126b4061b0530d124f3797426a08f72e90ef4994eeb348f5e92bd12d41cd3343a9e2"
           "71a2f73d2cc7ffbd65bf64fb63e759f312e615aae01ae9f4573a21f1a70f56a61cfbb94d"
           "8f96fcf06c2b3216ed9574f6888df86cd5e471b641507ac6815ca781f6d31e69d6848e54"
           "2a7c57dc21109b5574b63365a19273783fafc93639c414b9475ea5ea82e73958ff5fdba9"
           "67d52721ff71209e5a3db3c580e1bfd142ba4b8ab77eb16cb488d46a04a672662cd108b7"
           "e9c58ba13dfb850653208f81956539475ffce85e0b0da59e5bd8d90051be9b2cc99e37c0"
           "60ce09814e1524458bfb5427d7a16b672682be448fa16464fcb3e7f1dca6812a2c5a9814"
           "b98ccb676367b7b3b269c670cd0210edf70ad9cb337f766af75fe06d18b3f7f7c2eae656"
           "5ff2815c2c09b1a1f5", 16),
    c2=int("61803645f2798c06f2c08fc254eee612c55542051c8777d6ce69ede9c84a179afb2081"
           "167494dee727488ae5e9b56d98f4fcf132514616859fc854fbd3acf6aecd97324ac3f2af"
           "fa9f44864a9afc505754aa3b564b4617e887d6aa1f88095bccf6b47f458566f9d85e80fc"
           "d478a58d4c2e895d0ed428aa8919d8ce752472bdc704fe9f01b1f663e3a9defca4b38471"
           "34883d5433b6bebb7d5a0358bcc8e3385cdf8787a1c78165eb03fc295c2ee93809d7a7a4"
           "689e79faf173e4ca3d0a6a9175887d0c70b35c529aa02699c4d4e8c98a9f3b8f2be41f35"
           "905adebf8a6940a93875d1e24e578a93bdb7cbf66cd3cdb736466588649ac237d55121ce"
           "0c0d18bc5da660d8faf9f0849ed1775ffcc5edb6900ebfb6c1e33459d29655edf706324c"
           "f642c8f36433d6b850a43ee0e788e120737b8a2858d1b5302bad3413102fd7dccfe458b2"
           "57fdbf920fe942e23ec446b1b302d41710fe56b26e11987ac06cfa635664c7a0ec18f8c8"
           "c871919fc893a3117ff5e73d4c115e66e3bc5bd2b9127b2bb816c549245c65cf22a533a3"
           "d2b6cb7c46757d3a87173f93e8b431891697f8d60c59631734f46cf3d70d9065f0167d5a"
           "d7353c0812af024ced593273551d29c89232f2f3d548b9248291c1b8e833ed178eb2cf1a"
           "d6f1d6864f1fd3e2e3937e00d391ad330b443aec85528571740ed5538188c32caab27c7b"
           "f437df2bb97cb90e02", 16),
    c1c2=int("309f6e614d875e3bb0a77eedeb8895e7c6f297f161f576aef4f8b72bb5b81ef78b831a"
             "af134b09fe8697159cfd678c49920cb790e36580c5201a96848d7242fceb025808dd26b5"
             "0ff573ffca3f65e51b3b9fe85c7e44f5c8df0a9e524f64a5acc5c62cba7475978eb55e08"
             "93eff1c40547ef9db087f8a54a13bf33a4648c4719233cfb107ba469c61f1c07578d9c19"
             "fa8012b743d31fbca8eb4250ad902cf0c3d24c619fcd0874ad6a12ab8eafffabca6ed1aa"
             "a4ba0df1544c3826364ac955c5853dc0490b9992e867e2dc95ec4b8742f177b7b24f29f6"
             "8de4d552f32ca0da7d5cb2d85f020eefb8b58261c93643a4b63a9223efea803367b932b4"
             "30ae47730d9b493e4194cbc7e8aa6d8aae45aa016d7f197dab5bb9508d5af6c3f47c0ec4"
             "8ff604e53edbafa9a1bdae6add7169b83278a025f0be7980688806deaa9afaf80ca4212d"
             "53079c4841546bc1622c5bf211a9db1f8933211b6a5b5f312d6919181bf7797188645052"
             "a9fff167c7acbc43454cd3caab36a501feba27f28720f2ab23d5dea3c73d4421b059eef9"
             "f1c227a3ed59c487c9483a08e98bfd34920349fa861b41ce61a4caa8b7f0fc1fcba7dedb"
             "8f9c64ab3a42968f6c88f45541c734d7c0206968a103d02985854a5156d9edb99a332de9"
             "a6d47f9af6e68e18960fa5916cc48994334354d6303312b8e96602766bec337a8a92c596"
             "b21b6038828a6c9744", 16),
    m1m2=int("616263646566676869606a6b6c6d6e6f", 16),
    r0=int("57fb19590c31dc7c034b2a889cf4037ce3db799909c1eb0adb6199d8e96791daca9018"
           "891f34309daff32dced4af7d793d16734d055e28023acab7295956bfbfdf62bf0ccb2ed3"
           "1d5d176ca8b404e93007565fb6b72c33a512b4dc4f719231d62e27e34c3733929af32247"
           "f88c20d1ee77096cc80d3d642464054c815b35878ba812349c8bdc3c6b645daf1a0de609"
           "65f44dcf705681032480f1eeba82243196b96903becdc0df0801d4120cbd6db1c4b2841a"
           "27991c44a43750c24ed0825718ad14cfb9c6b40b78ff3d25f71741f2def1c9d420d4b0fa"
           "1e0a02e7851b5ec6a81133a368b80d1500b0f28fc653d2e6ff4366236dbf80ae3b4beae3"
           "5e04579f2c", 16),
    r1=int("6ee8ed76227672a7bcaa1e7f152c2ea39f2fa225f0713f58210c59b2270b110e38b650"
           "69aaedbeffc713c021336cc12f65227cc0357ca531c07c706e7224c2c11c3145bc0a05b1"
           "64f426ec03350820f9f416377e8720ddb577843cae929178bfe5772e2cc1e9b94e8fce81"
           "4eaf136c6ed218ca7b10ea4d5218e7ba82bd74bb9f19d3ccc7d2e140e91cfb25f76f54aa"
           "70f2ed88ef343dd5fb98617c0036b7717f7458ec847d7b52e8764a4e92c397133a95e35e"
           "9a82d5dc264ff423398cfadfbaec4727854e68f2e9e210d6a65c39b5a9b2a0ebdc538983"
           "4883680e42b5d8582344e3e07a01fbd6c46328dcfa03074d0bc02927f58466c2fa74ab60"
           "8177e3ec1b", 16),
    m0=int("414243444546474849404a4b4c4d4e4f", 16),
    m1=int("20202020202020202020202020202020", 16),
    # benchmark/bench_cryptography.cpp:49-63 (a valid DJN hs for the ISO key)
    hs=int("7788f6e8f57d3488cf9e0c7f4c19521de9aa172bf35924c7827a1189d6c688ac078f77"
           "7efcfc230e34f1fa5ae8d9d2ed5b062257618e0a0a485b0084b3fd39080031ea739bb48c"
           "dcce4ad41704ed930d40f53a1cc5d7f70bcb379f17a912b0ad14fabe8fc10213dcd1eabd"
           "9175ee9bf66c31e9af9703c9d92fa5c8d36279459631ba7e9d4571a10960f8e8d031b267"
           "22f6ae6f618895b9ce4fce926c8f54169168f6bb3e033861e08c2eca2161198481bc7c52"
           "3a38310be22f4dd7d028dc6b774e5cb8e6f33b24168697743b7deff411510e27694bf2e8"
           "0258b325fd97370f5110f54d8d7580b45ae3db26da4e3b0409f0cfbc56d9d9856b66d8bf"
           "46e727dc3148f70362d05faea743621e3841c94c78d53ee7e7fdef61022dd56922368991"
           "f843ca0aebf8436e5ec7e737c7ce72ac58f138bb11a3035fe96cc5a7b1aa9d565cb8a317"
           "f42564482dd3c842c5ee9fb523c165a8507ecee1ac4f185bdbcb7a51095c4c46bfe15aec"
           "3dfd77e1fd2b0003596df83bbb0d5521f16e2301ec2d4aafe25e4479ee965d8bb30a689a"
           "6f38ba710222fff7cf359d0f317b8e268f40f576c04262a595cdfc9a07b72978b9564ace"
           "699208291da7024e86b6eeb1458658852f10794c677b53db8577af272233722ad4579d7a"
           "074e57217e1c57d11862f74486c7f2987e4d09cd6fb2923569b577de50e89e6965a27e18"
           "7a8a341a7282b385ef", 16),
)


def enc(n, m, r, hs=None):
    """ipcl/pub_key.cpp:99-110 + :82-90 (non-DJN: r^n, DJN: hs^r)."""
    nsq = n * n
    obf = pow(hs, r, nsq) if hs is not None else pow(r, n, nsq)
    return ((n * m + 1) % nsq) * obf % nsq


def dec_crt(p, q, c):
    """ipcl/pri_key.cpp:114-167."""
    if q < p:
        p, q = q, p
    n = p * q
    g = n + 1
    hp = pow((pow(g % (p * p), p - 1, p * p) - 1) // p, -1, p)
    hq = pow((pow(g % (q * q), q - 1, q * q) - 1) // q, -1, q)
    pinv = pow(p, -1, q)
    mp = (pow(c % (p * p), p - 1, p * p) - 1) // p * hp % p
    mq = (pow(c % (q * q), q - 1, q * q) - 1) // q * hq % q
    return mp + ((mq - mp) * pinv % q) * p


def dec_raw(p, q, c):
    """ipcl/pri_key.cpp:92-111."""
    n = p * q
    lam = math.lcm(p - 1, q - 1)
    x = pow((pow(n + 1, lam, n * n) - 1) // n, -1, n)
    return (pow(c, lam, n * n) - 1) // n * x % n


def is_prime(n, rnd, rounds=24):
    if n < 2:
        return False
    for sp in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % sp == 0:
            return n == sp
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for _ in range(rounds):
        a = rnd.randrange(2, n - 1)
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def djn_keypair(bits, rnd):
    """p = q = 3 mod 4, gcd(p-1, q-1) = 2, |n| == bits (ipcl/keygen.cpp:73-90)."""
    half = bits // 2
    while True:
        def prime():
            while True:
                c = rnd.getrandbits(half) | (1 << (half - 1)) | (1 << (half - 2)) | 3
                if is_prime(c, rnd):
                    return c
        p, q = prime(), prime()
        if p == q or math.gcd(p - 1, q - 1) != 2:
            continue
        if (p * q).bit_length() != bits:
            continue
        return p, q


def djn_hs(n, rnd):
    """ipcl/pub_key.cpp:32-49: hs = (-x^2)^n mod n^2 for random x coprime to n."""
    while True:
        x = rnd.getrandbits(n.bit_length() + 128)
        if math.gcd(x, n) == 1:
            break
    h = (-(x % n) ** 2) % n
    return pow(h, n, n * n)


def hx(v):
    return format(v, "x")


def main():
    I = ISO
    n = I["p"] * I["q"]
    # pin to the reference's published answers before emitting anything
    assert enc(n, I["m0"], I["r0"]) == I["c1"]
    assert enc(n, I["m1"], I["r1"]) == I["c2"]
    assert I["c1"] * I["c2"] % (n * n) == I["c1c2"]
    assert dec_crt(I["p"], I["q"], I["c1c2"]) == I["m1m2"] == I["m0"] + I["m1"]
    assert dec_raw(I["p"], I["q"], I["c1c2"]) == I["m1m2"]
    lam = math.lcm(I["p"] - 1, I["q"] - 1)
    assert pow(I["hs"], lam, n * n) == 1  # hs is an n-th residue: valid DJN hs
    with open(os.path.join(HERE, "iso_18033_6.json"), "w") as f:
        json.dump({k: hx(v) for k, v in I.items()}, f, indent=1)

    rnd = random.Random(0xB200)
    # deterministic DJN keys for the other widths
    keys = {}
    for bits in (1024, 3072):
        p, q = djn_keypair(bits, rnd)
        nn = p * q
        keys[str(bits)] = dict(p=hx(p), q=hx(q), hs=hx(djn_hs(nn, rnd)))
    keys["2048"] = dict(p=hx(I["p"]), q=hx(I["q"]), hs=hx(I["hs"]))
    with open(os.path.join(HERE, "keys.json"), "w") as f:
        json.dump(keys, f, indent=1)

    # raw modexp vectors: random + edge cases per modulus width
    vec = []
    for bits in (512, 1024, 1536, 2048, 3072, 4096, 6144):
        R = 1 << bits
        mods = [rnd.getrandbits(bits) | (1 << (bits - 1)) | 1,   # top bit set
                rnd.getrandbits(bits - 37) | 1,                   # short modulus
                R - 1]                                            # all ones
        for m in mods:
            cases = [
                (rnd.randrange(m), rnd.getrandbits(bits)),
                (rnd.randrange(m), rnd.getrandbits(bits // 2)),
                (0, rnd.getrandbits(64)),
                (1, rnd.getrandbits(64)),
                (m - 1, rnd.getrandbits(64) | 1),
                (m - 1, rnd.getrandbits(64) & ~1),
                (rnd.randrange(m), 0),
                (rnd.randrange(m), 1),
                (rnd.randrange(m), 2),
                (rnd.randrange(m), (1 << 70)),        # top window zero-ish
                (rnd.randrange(m), 0xFFFFFFFF),
                (rnd.randrange(m, R), rnd.getrandbits(96)),  # base >= modulus
            ]
            for b, e in cases:
                vec.append(dict(bits=bits, b=hx(b), e=hx(e), m=hx(m),
                                r=hx(pow(b, e, m))))
    vec.append(dict(bits=512, b="5", e="3", m="1", r="0"))
    with open(os.path.join(HERE, "modexp_vectors.json"), "w") as f:
        json.dump(vec, f)

    # scheme vectors per key width: encrypt (DJN and not), decrypt, add, mul
    scheme = {}
    for bits, kd in keys.items():
        p, q, hs = int(kd["p"], 16), int(kd["q"], 16), int(kd["hs"], 16)
        nn = p * q
        nsq = nn * nn
        items = []
        for i in range(6):
            m = [0, 1, nn - 1, rnd.randrange(nn), rnd.getrandbits(32),
                 rnd.randrange(nn)][i]
            r_djn = rnd.getrandbits(int(bits) // 2)
            r_std = rnd.randrange(1, nn)
            c_djn = enc(nn, m, r_djn, hs)
            c_std = enc(nn, m, r_std)
            assert dec_crt(p, q, c_djn) == m and dec_crt(p, q, c_std) == m
            assert dec_raw(p, q, c_djn) == m
            k = rnd.getrandbits(32) if i % 2 else rnd.randrange(nn)
            items.append(dict(m=hx(m), r_djn=hx(r_djn), r_std=hx(r_std),
                              c_djn=hx(c_djn), c_std=hx(c_std),
                              c_plain=hx((nn * m + 1) % nsq),
                              k=hx(k), c_mul=hx(pow(c_djn, k, nsq)),
                              c_add=hx(c_djn * c_std % nsq)))
        scheme[bits] = items
    with open(os.path.join(HERE, "scheme_vectors.json"), "w") as f:
        json.dump(scheme, f)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
