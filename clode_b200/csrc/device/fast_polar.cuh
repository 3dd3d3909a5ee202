// fast_polar.cuh — the scale factor of the Marsaglia polar method, sqrt(-2 log(q) / q), for production double builds
// of the stochastic stepper (not the bit-exact tier, not single precision).
//
// The reference computes it as written (clODE_random.cl:107): a library log (libdevice: ~25 FP64-pipe instructions and
// a branch), an IEEE division (8 + a slow-path branch) and an IEEE square root (~9 + a slow-path call) per pair of
// variates — a third of the instructions of a stochastic-Euler step of the lactotroph model (C4).  Here:
//   * s = -2 log q, q in (0, 1), from a 256-entry table: q = 2^e' m with m in [0.75, 1.5) (the upper half of a binade is
//     folded down, so that q -> 1 means e' = 0 and m -> 1: no cancellation against e' ln2), j = top 8 mantissa bits,
//     r = m rc_j - 1 (|r| <= 2^-9; one FMA), s = e' (-2 ln2) + (-2 log c_j) - 2 log1p(r), the first two terms as hi + lo
//     pairs, log1p by a degree-6 polynomial: 12 FP64 instructions, no branch.  The last interval (m in [1 - 1/512, 1))
//     has rc = 1 and log c = 0, so next to q = 1 the result keeps its full RELATIVE accuracy;
//   * sqrt(s / q) = s rsqrt(s q): one SFU seed (MUFU.RSQ64H, rsqrt.approx.ftz.f64) and ONE third-order correction,
//     y (1 + e/2 + 3 e^2/8) with e = 1 - s q y^2 (|e| < 2^-19: the remainder is below 2^-58): 7 FP64 instructions instead
//     of the division and the square root.
// Error of the factor: <= 3 ulp against 80-bit arithmetic with a pessimistic model of the seed (2.6 measured; an OpenCL
// device may return a log that is 3 ulp off, i.e. the same for the composite; tests/emu/fast_polar_check.cpp, driven by
// tests/test_fast_exp.py; the GPU test compares the variates of a C4-like run with the oracle's).  The integer RNG
// stream is untouched: the accept / reject test of the pair stays where it is (rng.cuh), in contraction-proof arithmetic.
#ifndef CLODE_FAST_POLAR_CUH
#define CLODE_FAST_POLAR_CUH

struct ClodePolarEntry { double rc, hi, lo; }; // 1 / c_j rounded, -2 log(c_j) = 2 log(rc) as hi + lo   (scripts/make_polar_table.py)
__device__ const ClodePolarEntry clode_polar_table[256] = {
    {0x1.ff007fc01ff00p-1, -0x1.ff802a9ab11e6p-9, -0x1.e29e3a153e432p-63}, {0x1.fd04794a10e6ap-1, -0x1.7ee11ebd82ec4p-7, -0x1.3c2d23a074505p-62},
    {0x1.fb0c610d5e939p-1, -0x1.3e7295d25a7d5p-6, -0x1.600d65eebbc60p-60}, {0x1.f9182b6813bafp-1, -0x1.bcf712c743853p-6, 0x1.7b4213447a4ccp-60},
    {0x1.f727cce5f530ap-1, -0x1.1d7f7eb9eebf1p-5, -0x1.2be019c2d240ep-60}, {0x1.f53b3a3fa204ep-1, -0x1.5c45a51b8d393p-5, 0x1.885b61f610d84p-62},
    {0x1.f3526859b8cecp-1, -0x1.9ace7551cc515p-5, 0x1.cbf63e207e981p-59}, {0x1.f16d4c4401f17p-1, -0x1.d91a66c543cbep-5, -0x1.4b2c67dcd0956p-59},
    {0x1.ef8bdb389ebadp-1, -0x1.0b94f7c196173p-4, 0x1.c5bc089e0b23bp-58}, {0x1.edae0a9b3d3a5p-1, -0x1.2a7ec2214e879p-4, -0x1.042b74e00f373p-59},
    {0x1.ebd3cff850b0cp-1, -0x1.494acc34d911dp-4, 0x1.9d6a40b6e333bp-58}, {0x1.e9fd21044e799p-1, -0x1.67f94f094bd92p-4, -0x1.19f3f276b596dp-58},
    {0x1.e829f39aef509p-1, -0x1.868a83083f6d0p-4, 0x1.284d2b1a4a1edp-59}, {0x1.e65a3dbe74d6bp-1, -0x1.a4fe9ffa3d233p-4, 0x1.4502014926cd6p-58},
    {0x1.e48df596f3394p-1, -0x1.c355dd0921f2fp-4, 0x1.4d9501f1df1d3p-58}, {0x1.e2c511719ee16p-1, -0x1.e19070c276010p-4, -0x1.a66e7585e8241p-58},
    {0x1.e0ff87c01e100p-1, -0x1.ffae9119b92fbp-4, -0x1.ba13162a9c44ep-59}, {0x1.df3d4f17de4dbp-1, -0x1.0ed839b5526fep-3, -0x1.1e4add513114dp-57},
    {0x1.dd7e5e316d94cp-1, -0x1.1dcb263db1944p-3, -0x1.7d7a7a2605718p-57}, {0x1.dbc2abe7d71d4p-1, -0x1.2cb0283f5de22p-3, 0x1.34d66a3f7a2b6p-57},
    {0x1.da0a2f3803b41p-1, -0x1.3b87598b1b6f0p-3, -0x1.44d6a6b9dad0cp-57}, {0x1.d854df401d855p-1, -0x1.4a50d3aa1b03fp-3, 0x1.9308973e22a83p-60},
    {0x1.d6a2b33ef7448p-1, -0x1.590cafdf01c26p-3, 0x1.42a375515892ep-57}, {0x1.d4f3a293769cap-1, -0x1.67bb0726ec0fbp-3, -0x1.d2da7bd644829p-58},
    {0x1.d347a4bc01d34p-1, -0x1.765bf23a6be17p-3, -0x1.cff28ef6a5931p-57}, {0x1.d19eb155f08a4p-1, -0x1.84ef898e82828p-3, 0x1.f491a5df236ecp-57},
    {0x1.cff8c01cff8c0p-1, -0x1.9375e55595edfp-3, 0x1.e463f9e4dd91fp-58}, {0x1.ce55c8eac7900p-1, -0x1.a1ef1d8061cd8p-3, -0x1.76df97bcb1787p-59},
    {0x1.ccb5c3b636e3ap-1, -0x1.b05b49bee4403p-3, 0x1.89f383dad0d65p-57}, {0x1.cb18a8930de60p-1, -0x1.beba818146764p-3, -0x1.d248382a5ecffp-61},
    {0x1.c97e6fb15e44dp-1, -0x1.cd0cdbf8c13e0p-3, 0x1.64af228bcf63ap-59}, {0x1.c7e7115d0ce95p-1, -0x1.db5270187d925p-3, 0x1.9d4a8f3f05aa5p-58},
    {0x1.c65285fd56843p-1, -0x1.e98b54967146bp-3, -0x1.a227143a5a99ap-57}, {0x1.c4c0c61456a8ep-1, -0x1.f7b79fec37de2p-3, -0x1.38176812fe880p-58},
    {0x1.c331ca3e91679p-1, -0x1.02ebb42bf3d4ap-2, 0x1.652e70072e4b1p-56}, {0x1.c1a58b327f576p-1, -0x1.09f561ee719c4p-2, 0x1.aae2afa34f48ap-57},
    {0x1.c01c01c01c01cp-1, -0x1.10f8e422539b1p-2, -0x1.cf798d39f1b7dp-57}, {0x1.be9526d0769fap-1, -0x1.17f6458fca611p-2, 0x1.f52f6c3723f80p-56},
    {0x1.bd10f365451b6p-1, -0x1.1eed90e2dc2c3p-2, -0x1.837097648f581p-57}, {0x1.bb8f609879493p-1, -0x1.25ded0abc6ad3p-2, -0x1.14f176448b993p-59},
    {0x1.ba10679bd8488p-1, -0x1.2cca0f5f5f252p-2, -0x1.dcdca01dc0febp-56}, {0x1.b89401b89401cp-1, -0x1.33af575770e4dp-2, -0x1.f28bf9ca923d6p-57},
    {0x1.b71a284ee6b34p-1, -0x1.3a8eb2d31a375p-2, -0x1.bbbeaea81ece2p-56}, {0x1.b5a2d4d5b081fp-1, -0x1.41682bf727bbfp-2, 0x1.1e103f093930dp-57},
    {0x1.b42e00da17007p-1, -0x1.483bccce6e3dcp-2, -0x1.b1391fb1b4b22p-56}, {0x1.b2bba5ff26a23p-1, -0x1.4f099f4a230b1p-2, -0x1.24140543648f3p-57},
    {0x1.b14bbdfd760e6p-1, -0x1.55d1ad4232d70p-2, 0x1.4644b3703041cp-56}, {0x1.afde42a2cb482p-1, -0x1.5c940075972b9p-2, 0x1.1919a4664319dp-56},
    {0x1.ae732dd1c2a09p-1, -0x1.6350a28aaa759p-2, 0x1.0ea8fd38a2c66p-57}, {0x1.ad0a798177693p-1, -0x1.6a079d0f7aad0p-2, -0x1.28891a29eac08p-56},
    {0x1.aba41fbd2e5b1p-1, -0x1.70b8f97a1aa74p-2, 0x1.de12ad4822814p-56}, {0x1.aa401aa401aa4p-1, -0x1.7764c128f2127p-2, -0x1.440d1e78f44cep-56},
    {0x1.a8de64688ebabp-1, -0x1.7e0afd630c276p-2, 0x1.d9f13877e61b9p-56}, {0x1.a77ef750a56dap-1, -0x1.84abb75865137p-2, 0x1.16fa715e8d38bp-58},
    {0x1.a621cdb4f8fdfp-1, -0x1.8b46f8223625bp-2, -0x1.610816ebe4976p-56}, {0x1.a4c6e200d2637p-1, -0x1.91dcc8c340bdfp-2, 0x1.f28442017473fp-56},
    {0x1.a36e2eb1c432dp-1, -0x1.986d3228180c8p-2, -0x1.0593750fffe78p-57}, {0x1.a217ae575ff2fp-1, -0x1.9ef83d2769a34p-2, 0x1.9fb3f9cdff9d3p-56},
    {0x1.a0c35b92ecdf1p-1, -0x1.a57df28244dcbp-2, 0x1.966bc4ca8938dp-56}, {0x1.9f713117200d0p-1, -0x1.abfe5ae46124ap-2, -0x1.2b1a83b18de21p-57},
    {0x1.9e2129a7d5f0ap-1, -0x1.b2797ee46320cp-2, -0x1.1adf25feae309p-56}, {0x1.9cd34019cd340p-1, -0x1.b8ef670420c3bp-2, -0x1.9990bc47005e0p-58},
    {0x1.9b876f5262dd1p-1, -0x1.bf601bb0e44e0p-2, 0x1.beb83c874aaf3p-56}, {0x1.9a3db2474fb98p-1, -0x1.c5cba543ae424p-2, 0x1.44269756071afp-57},
    {0x1.98f603fe670a0p-1, -0x1.cc320c0176501p-2, -0x1.cd329bc9d42b1p-63}, {0x1.97b05f8d56652p-1, -0x1.d293581b6b3e7p-2, 0x1.204a2aa97ac8ep-57},
    {0x1.966cc01966cc0p-1, -0x1.d8ef91af31d5ep-2, -0x1.d01e4d9c3e3a7p-56}, {0x1.952b20d73ee97p-1, -0x1.df46c0c722d30p-2, 0x1.4f486fc6e8c8ap-63},
    {0x1.93eb7d0aa6759p-1, -0x1.e598ed5a87e2ep-2, 0x1.daf3c7a62832cp-56}, {0x1.92add0064ab74p-1, -0x1.ebe61f4dd7b0bp-2, 0x1.9987ee52650b9p-59},
    {0x1.9172152b841ddp-1, -0x1.f22e5e72f105cp-2, 0x1.98a0bf20f9d99p-58}, {0x1.903847ea1cec1p-1, -0x1.f871b28955045p-2, -0x1.8d2b5b2204b4cp-56},
    {0x1.8f0063c018f00p-1, -0x1.feb0233e607cep-2, -0x1.6e32d5e8c7080p-56}, {0x1.8dca64397e408p-1, -0x1.0274dc16c232fp-1, 0x1.6bb183e51ec40p-55},
    {0x1.8c9644f01efbcp-1, -0x1.058f3c703ebc5p-1, -0x1.e9432dc9528f1p-55}, {0x1.8b64018b64019p-1, -0x1.08a73667c57aep-1, -0x1.2140c5a328e6dp-55},
    {0x1.8a3395c018a34p-1, -0x1.0bbccdb0d24bcp-1, 0x1.2333a23204a40p-55}, {0x1.8904fd503744bp-1, -0x1.0ed005f657da5p-1, -0x1.0b5e955ff414ep-58},
    {0x1.87d8340ab6e97p-1, -0x1.11e0e2dad9cb6p-1, -0x1.97b8198d22e05p-55}, {0x1.86ad35cb59a84p-1, -0x1.14ef67f88685ap-1, -0x1.a6880da1b13e4p-57},
    {0x1.8583fe7a7c018p-1, -0x1.17fb98e15095ep-1, -0x1.1458b5d97ba9dp-55}, {0x1.845c8a0ce5129p-1, -0x1.1b05791f07b4ap-1, 0x1.b26dc55e2d052p-55},
    {0x1.8336d48397a24p-1, -0x1.1e0d0c33716bdp-1, -0x1.154d86a4ff98bp-58}, {0x1.8212d9eba4018p-1, -0x1.211255986160cp-1, 0x1.3a2eb579e2857p-58},
    {0x1.80f0965dfabcbp-1, -0x1.241558bfd1405p-1, 0x1.99bae06a5c863p-60}, {0x1.7fd005ff40180p-1, -0x1.27161913f853dp-1, 0x1.0e09ea9b4c4a4p-55},
    {0x1.7eb124ffa053bp-1, -0x1.2a1499f762bcap-1, -0x1.895c18aa47a54p-56}, {0x1.7d93ef9aa4b46p-1, -0x1.2d10dec508582p-1, -0x1.f3ee1106a6ca7p-56},
    {0x1.7c7862170949fp-1, -0x1.300aead06350cp-1, 0x1.95d2280d51407p-57}, {0x1.7b5e78c693733p-1, -0x1.3302c1658658ap-1, 0x1.263d5c1f755e9p-55},
    {0x1.7a463005e918cp-1, -0x1.35f865c93293ep-1, -0x1.8d8af2d5b0557p-58}, {0x1.792f843c689c3p-1, -0x1.38ebdb38ed320p-1, -0x1.2d733ea6502f0p-55},
    {0x1.781a71dc01782p-1, -0x1.3bdd24eb14b69p-1, -0x1.06d1e3224d3e9p-56}, {0x1.7706f5610d8d0p-1, -0x1.3ecc460ef5f50p-1, 0x1.0c4f82601ebfap-59},
    {0x1.75f50b522b17cp-1, -0x1.41b941cce0beep-1, 0x1.8027c87f91214p-56}, {0x1.74e4b040174e5p-1, -0x1.44a41b463c47bp-1, 0x1.430c8309edcfcp-55},
    {0x1.73d5e0c5899f7p-1, -0x1.478cd5959b3d8p-1, 0x1.1c0f372f6825bp-56}, {0x1.72c899870f91fp-1, -0x1.4a7373cecf997p-1, 0x1.51d7e6a892849p-56},
    {0x1.71bcd732e940ap-1, -0x1.4d57f8fefe27fp-1, -0x1.cb3fe83434321p-55}, {0x1.70b29680e66fap-1, -0x1.503a682cb1cb3p-1, 0x1.bc78b7cdae677p-55},
    {0x1.6fa9d43244380p-1, -0x1.531ac457ee77fp-1, 0x1.c4826ceaff1c8p-55}, {0x1.6ea28d118b474p-1, -0x1.55f9107a43ee2p-1, 0x1.81de37d2989eep-55},
    {0x1.6d9cbdf26eaefp-1, -0x1.58d54f86e02f3p-1, 0x1.24f586adeb499p-56}, {0x1.6c9863b1ab429p-1, -0x1.5baf846aa1b1ap-1, -0x1.ec1e3016fc9f5p-57},
    {0x1.6b957b34e7803p-1, -0x1.5e87b20c2954ap-1, 0x1.fa7088c705f8ap-55}, {0x1.6a94016a94017p-1, -0x1.615ddb4bec13cp-1, 0x1.e15bd0fed391dp-55},
    {0x1.6993f349cc726p-1, -0x1.64320304447c1p-1, 0x1.d617f8a08338cp-57}, {0x1.68954dd2390bap-1, -0x1.67042c0983e30p-1, -0x1.b9b7b9e219186p-55},
    {0x1.67980e0bf08c7p-1, -0x1.69d4592a0362ep-1, 0x1.fc80d000b4083p-56}, {0x1.669c31075ab40p-1, -0x1.6ca28d2e34986p-1, 0x1.5e8e76dd346a0p-55},
    {0x1.65a1b3dd13357p-1, -0x1.6f6ecad8b2292p-1, 0x1.fc083df227104p-58}, {0x1.64a893adcd25fp-1, -0x1.723914e6500e2p-1, -0x1.4caf721f626aap-56},
    {0x1.63b0cda236e1cp-1, -0x1.75016e0e2ba63p-1, -0x1.a748662fc4171p-55}, {0x1.62ba5eeade65ep-1, -0x1.77c7d901bb913p-1, 0x1.29943804dfbeep-55},
    {0x1.61c544c0161c5p-1, -0x1.7a8c586cdf545p-1, 0x1.9a576c0601322p-57}, {0x1.60d17c61da198p-1, -0x1.7d4eeef5eec6ep-1, 0x1.58f8f27d8e90fp-56},
    {0x1.5fdf0317b5c6fp-1, -0x1.800f9f3dc94ccp-1, 0x1.306488dd76781p-57}, {0x1.5eedd630a9fb3p-1, -0x1.82ce6bdfe4d9ep-1, 0x1.dc45997fbc413p-55},
    {0x1.5dfdf303137b6p-1, -0x1.858b57725cc43p-1, -0x1.7cab36811fa33p-56}, {0x1.5d0f56ec91e57p-1, -0x1.8846648600623p-1, 0x1.f4419b612c65ap-56},
    {0x1.5c21ff51ef005p-1, -0x1.8aff95a661781p-1, -0x1.9f4fca257a85dp-56}, {0x1.5b35e99f06714p-1, -0x1.8db6ed59e272dp-1, 0x1.51ec4c14526a6p-55},
    {0x1.5a4b1346add2bp-1, -0x1.906c6e21c4753p-1, -0x1.dd8e962c0c0adp-55}, {0x1.596179c29d2cep-1, -0x1.93201a7a35336p-1, 0x1.02711f5645823p-56},
    {0x1.58791a9357ccep-1, -0x1.95d1f4da5ca0ap-1, 0x1.f3c3fbbc738aap-56}, {0x1.5791f34015792p-1, -0x1.9881ffb46a6f0p-1, 0x1.951ec6ae7473ep-57},
    {0x1.56ac0156ac015p-1, -0x1.9b303d75a3620p-1, -0x1.6ef49cf67f73bp-55}, {0x1.55c7426b79286p-1, -0x1.9ddcb0866e742p-1, -0x1.0f947c24d6d15p-56},
    {0x1.54e3b4194ce66p+0, 0x1.25410494e56c8p-1, -0x1.da7e21101b5adp-56}, {0x1.5401540154015p+0, 0x1.22981fbef797ap-1, 0x1.b53ed4fe4c507p-56},
    {0x1.53201fcb02fb1p+0, 0x1.1ff0fe7cf47a9p-1, -0x1.a15d801e7d762p-56}, {0x1.5240152401524p+0, 0x1.1d4b9e796c245p-1, 0x1.233e2172b6715p-55},
    {0x1.516131c015161p+0, 0x1.1aa7fd638d33ep-1, 0x1.529616f79ff4ep-56}, {0x1.508373590ec9cp+0, 0x1.180618ef18adep-1, -0x1.7e4369c72b404p-58},
    {0x1.4fa6d7aeb597cp+0, 0x1.1565eed455fc2p-1, 0x1.829024aa2ed78p-55}, {0x1.4ecb5c86b3d24p+0, 0x1.12c77cd00713cp-1, 0x1.1522847de5d12p-55},
    {0x1.4df0ffac83c01p+0, 0x1.102ac0a35cc1bp-1, 0x1.94404052f3458p-57}, {0x1.4d17bef15cb4ep+0, 0x1.0d8fb813eb1efp-1, 0x1.5a21d4fe8d42ap-55},
    {0x1.4c3f982c20723p+0, 0x1.0af660eb9e278p-1, -0x1.440ad727f641bp-56}, {0x1.4b68893948d1cp+0, 0x1.085eb8f8ae799p-1, -0x1.3d8174030ad14p-56},
    {0x1.4a928ffad5b5cp+0, 0x1.05c8be0d9635ap-1, 0x1.a38ef996b0c96p-57}, {0x1.49bdaa583b401p+0, 0x1.03346e0106062p-1, -0x1.9475699c6a38ep-55},
    {0x1.48e9d63e504d1p+0, 0x1.00a1c6adda472p-1, 0x1.05a22e785ea23p-57}, {0x1.4817119f3d325p+0, 0x1.fc218be620a5fp-2, -0x1.be438c2581880p-57},
    {0x1.47455a726abf2p+0, 0x1.f702d36777df0p-2, 0x1.8ae998c1dd664p-57}, {0x1.4674aeb4717e9p+0, 0x1.f1e75fadf9bdep-2, 0x1.59b44f8126332p-57},
    {0x1.45a50c670938fp+0, 0x1.eccf2c8fe920bp-2, 0x1.217062a6fe69fp-57}, {0x1.44d67190f8b43p+0, 0x1.e7ba35eb77e2ap-2, 0x1.ec7721b26dd59p-56},
    {0x1.4408dc3e05b22p+0, 0x1.e2a877a6b2c0fp-2, -0x1.6d10f1efcca1bp-56}, {0x1.433c4a7ee52b4p+0, 0x1.dd99edaf6d7e9p-2, 0x1.4cb1c548a6ce6p-58},
    {0x1.4270ba692bc4dp+0, 0x1.d88e93fb2f451p-2, 0x1.f7fb96815e081p-56}, {0x1.41a62a173e821p+0, 0x1.d38666871f467p-2, -0x1.4b38932bc0bedp-59},
    {0x1.40dc97a843ae8p+0, 0x1.ce816157f1985p-2, -0x1.6ba2099514bdbp-56}, {0x1.4014014014014p+0, 0x1.c97f8079d44ecp-2, 0x1.41a8c6e6c4ee7p-56},
    {0x1.3f4c65072bf74p+0, 0x1.c480c0005cccfp-2, 0x1.49abc89ceca67p-56}, {0x1.3e85c12a9d651p+0, 0x1.bf851c067555cp-2, -0x1.c9302152b2212p-57},
    {0x1.3dc013dc013dcp+0, 0x1.ba8c90ae4ad19p-2, 0x1.afe88865b42bdp-56}, {0x1.3cfb5b51698ebp+0, 0x1.b5971a213acd9p-2, -0x1.35f155b885f1fp-57},
    {0x1.3c3795c553afbp+0, 0x1.b0a4b48fc1b44p-2, -0x1.6ab87331d9cbfp-57}, {0x1.3b74c1769aa5cp+0, 0x1.abb55c31693aep-2, 0x1.a9a875993ea8ap-58},
    {0x1.3ab2dca869b81p+0, 0x1.a6c90d44b704cp-2, -0x1.67e06f618b545p-56}, {0x1.39f1e5a22f36ep+0, 0x1.a1dfc40f1b7f1p-2, -0x1.ce009e6f018ffp-56},
    {0x1.3931daaf8f721p+0, 0x1.9cf97cdce0ec1p-2, -0x1.e779df58e47ddp-58}, {0x1.3872ba2057e04p+0, 0x1.981634011aa74p-2, -0x1.64c2df743bd5ap-56},
    {0x1.37b4824872744p+0, 0x1.9335e5d594985p-2, 0x1.d8757a8fb3347p-56}, {0x1.36f7317fd9212p+0, 0x1.8e588ebac2dc1p-2, 0x1.d2acb445001d8p-57},
    {0x1.363ac622898b1p+0, 0x1.897e2b17b19a6p-2, -0x1.4f380cbe9dbe8p-56}, {0x1.357f3e9078e5bp+0, 0x1.84a6b759f512dp-2, -0x1.6156fc3047cf8p-58},
    {0x1.34c4992d87fd9p+0, 0x1.7fd22ff599d4cp-2, -0x1.5bf457b7d1812p-57}, {0x1.340ad461776d3p+0, 0x1.7b0091651528bp-2, 0x1.10d3e606a318fp-57},
    {0x1.3351ee97dbfc6p+0, 0x1.7631d82935a84p-2, -0x1.8dc7c5f3e101cp-56}, {0x1.3299e6401329ap+0, 0x1.716600c914055p-2, 0x1.855f3b0e0e1cdp-58},
    {0x1.31e2b9cd37dc2p+0, 0x1.6c9d07d203fc4p-2, -0x1.fafd9b2dc9d46p-61}, {0x1.312c67b6173eep+0, 0x1.67d6e9d785770p-2, -0x1.0185383697ee2p-58},
    {0x1.3076ee7525c2cp+0, 0x1.6313a37335d76p-2, 0x1.cab0de1592fb0p-57}, {0x1.2fc24c8874486p+0, 0x1.5e533144c1718p-2, 0x1.b8189ade2b075p-56},
    {0x1.2f0e8071a5703p+0, 0x1.59958ff1d52f4p-2, -0x1.e65da72814af4p-57}, {0x1.2e5b88b5e3104p+0, 0x1.54dabc26105d3p-2, -0x1.42346e5e4fa23p-57},
    {0x1.2da963ddd3cfbp+0, 0x1.5022b292f6a45p-2, 0x1.0ff9b512dbc1dp-58}, {0x1.2cf8107590e67p+0, 0x1.4b6d6fefe22a5p-2, 0x1.fcf56e7951abbp-57},
    {0x1.2c478d0c9c013p+0, 0x1.46baf0f9f5db8p-2, 0x1.717c37bdf2e08p-56}, {0x1.2b97d835d548ep+0, 0x1.420b32740fdd6p-2, 0x1.8e9bd2fbbdd69p-56},
    {0x1.2ae8f087718d0p+0, 0x1.3d5e3126bc281p-2, -0x1.e83d7b49da757p-56}, {0x1.2a3ad49af0907p+0, 0x1.38b3e9e027477p-2, -0x1.98a8b82ff1eb3p-56},
    {0x1.298d830d13780p+0, 0x1.340c59741142dp-2, 0x1.18413163ccbcfp-58}, {0x1.28e0fa7dd35a3p+0, 0x1.2f677cbbc0a98p-2, 0x1.42160f40d56bbp-59},
    {0x1.2835399057efdp+0, 0x1.2ac55095f5c5bp-2, -0x1.2b68636453e34p-56}, {0x1.278a3eeaee650p+0, 0x1.2625d1e6ddf55p-2, 0x1.4e87b0e13f0a5p-58},
    {0x1.26e009370049cp+0, 0x1.2188fd9807266p-2, 0x1.a3015e71fdb2bp-56}, {0x1.263697210aa18p+0, 0x1.1ceed09853755p-2, 0x1.e3736a838a6b8p-62},
    {0x1.258de75895121p+0, 0x1.185747dbecf34p-2, -0x1.1ee90992dcbabp-57}, {0x1.24e5f89029305p+0, 0x1.13c2605c398bfp-2, 0x1.da26b09af7476p-56},
    {0x1.243ec97d49eaep+0, 0x1.0f301717cf0fbp-2, -0x1.f8835d0d8979fp-56}, {0x1.239858d86b11fp+0, 0x1.0aa06912675d5p-2, 0x1.68a3f37b5ce5ap-57},
    {0x1.22f2a55ce8fc5p+0, 0x1.06135354d4b19p-2, -0x1.575f2fc45ac69p-57}, {0x1.224dadc900489p+0, 0x1.0188d2ecf613ep-2, 0x1.451cff9dfe3fbp-58},
    {0x1.21a970ddc5ba7p+0, 0x1.fa01c9db57ce7p-3, 0x1.1c0b6eb19fd48p-59}, {0x1.2105ed5f1e336p+0, 0x1.f0f70cdd992e4p-3, 0x1.9db09cb07729cp-57},
    {0x1.20632213b6c6dp+0, 0x1.e7f1691a32d3ap-3, 0x1.7990e21019877p-57}, {0x1.1fc10dc4fce8bp+0, 0x1.def0d8d466dbbp-3, 0x1.0efb45962e028p-57},
    {0x1.1f1faf3f16b64p+0, 0x1.d5f55659210e1p-3, -0x1.b19f3d5cb5706p-58}, {0x1.1e7f0550db594p+0, 0x1.ccfedbfee13a8p-3, 0x1.32fe71255a574p-59},
    {0x1.1ddf0ecbcb841p+0, 0x1.c40d6425a5cb4p-3, 0x1.987464c3722b2p-57}, {0x1.1d3fca840a074p+0, 0x1.bb20e936d6976p-3, 0x1.f2ae991c88432p-61},
    {0x1.1ca13750547fep+0, 0x1.b23965a52ff04p-3, -0x1.e9dd426e0f27bp-57}, {0x1.1c035409fc1dfp+0, 0x1.a956d3ecade60p-3, -0x1.cacff4ed42aa4p-57},
    {0x1.1b661f8cde833p+0, 0x1.a0792e9277cadp-3, 0x1.fc9b2957205c6p-57}, {0x1.1ac998b75eb90p+0, 0x1.97a07024cbe6ep-3, -0x1.82e641279cfb5p-60},
    {0x1.1a2dbe6a5e3e4p+0, 0x1.8ecc933aeb6e2p-3, -0x1.9be67f7aa7546p-60}, {0x1.19928f89362b7p+0, 0x1.85fd927506a46p-3, -0x1.0665c3071db3dp-61},
    {0x1.18f80af9b06dcp+0, 0x1.7d33687c293c8p-3, -0x1.0f063e63e7076p-57}, {0x1.185e2fa401186p+0, 0x1.746e100226edbp-3, -0x1.4b70f10e93174p-58},
    {0x1.17c4fc72bfcb9p+0, 0x1.6bad83c1883bap-3, 0x1.ae60449356c12p-57}, {0x1.172c7052e1316p+0, 0x1.62f1be7d7774ap-3, -0x1.5fb58f1376e6ep-62},
    {0x1.16948a33b08fap+0, 0x1.5a3abb01ade21p-3, 0x1.e4f357d0bf567p-58}, {0x1.15fd4906c96f1p+0, 0x1.5188742261311p-3, 0x1.996258b3d8a77p-59},
    {0x1.1566abc011567p+0, 0x1.48dae4bc3101dp-3, 0x1.b90461005f525p-58}, {0x1.14d0b155b19aep+0, 0x1.403207b414b79p-3, 0x1.a95502af7fe71p-57},
    {0x1.143b58c01143bp+0, 0x1.378dd7f74970fp-3, -0x1.2d70e0535f54fp-59}, {0x1.13a6a0f9cf01ep+0, 0x1.2eee507b402ffp-3, -0x1.a1228837a052dp-58},
    {0x1.131288ffbb3b6p+0, 0x1.26536c3d8c36cp-3, -0x1.c9fb41d22e910p-57}, {0x1.127f0fd0d2295p+0, 0x1.1dbd2643d1913p-3, -0x1.fc9a20edb0203p-57},
    {0x1.11ec346e36092p+0, 0x1.152b799bb3cd0p-3, -0x1.e90703082910cp-58}, {0x1.1159f5db29606p+0, 0x1.0c9e615ac4e19p-3, -0x1.0fed164d13b5bp-57},
    {0x1.10c8531d0952ep+0, 0x1.0415d89e7444bp-3, 0x1.40b9e3aea6c39p-58}, {0x1.10374b3b480aap+0, 0x1.f723b517fc51fp-4, -0x1.c6eab08695901p-58},
    {0x1.0fa6dd3f67322p+0, 0x1.e624c4a0b5e15p-4, -0x1.a3a33b3446795p-58}, {0x1.0f170834f27fap+0, 0x1.d52ed6405d87ap-4, -0x1.4a8a6ef59ba39p-61},
    {0x1.0e87cb297a51ep+0, 0x1.c441e06f72a93p-4, -0x1.45b3d79755aa4p-58}, {0x1.0df9252c8e5e6p+0, 0x1.b35dd9b58baa8p-4, -0x1.94985538de795p-62},
    {0x1.0d6b154fb86f9p+0, 0x1.a282b8a936174p-4, -0x1.8c077e47149d6p-59}, {0x1.0cdd9aa677344p+0, 0x1.91b073efd7314p-4, -0x1.4fddb2a56c208p-63},
    {0x1.0c50b446391f3p+0, 0x1.80e7023d8ccc8p-4, -0x1.ab7945fa2720bp-58}, {0x1.0bc4614657569p+0, 0x1.70265a550e77bp-4, 0x1.e3b80a8c6332fp-58},
    {0x1.0b38a0c010b39p+0, 0x1.5f6e73078efc3p-4, 0x1.affdb6d68f1fbp-61}, {0x1.0aad71ce84d16p+0, 0x1.4ebf43349e26ap-4, 0x1.fc23106232514p-58},
    {0x1.0a22d38eaf2bfp+0, 0x1.3e18c1ca0ae99p-4, 0x1.27edc6f1c907ep-60}, {0x1.0998c51f624d5p+0, 0x1.2d7ae5c3c5bb7p-4, 0x1.15d312cc97c03p-58},
    {0x1.090f45a1430aap+0, 0x1.1ce5a62bc3540p-4, -0x1.839390333b61ep-58}, {0x1.08865436c3cf7p+0, 0x1.0c58fa19dfaabp-4, -0x1.62b162f225e0bp-59},
    {0x1.07fdf0041ff7cp+0, 0x1.f7a9b16782855p-5, 0x1.c938df3eb88aap-59}, {0x1.0776182f57386p+0, 0x1.d6b272597981fp-5, 0x1.95e5c8f8f355ep-60},
    {0x1.06eecbe029155p+0, 0x1.b5cc258b718e7p-5, -0x1.791d41005f9a7p-59}, {0x1.06680a4010668p+0, 0x1.94f6b99a24473p-5, -0x1.0693080ae9e8ap-63},
    {0x1.05e1d27a3ee9cp+0, 0x1.74321d3d006d2p-5, 0x1.690fe9477840cp-59}, {0x1.055c23bb98e2ap+0, 0x1.537e3f45f354ep-5, -0x1.b169406d66a7bp-59},
    {0x1.04d6fd32b0c7bp+0, 0x1.32db0ea132e10p-5, -0x1.e767bb50221ffp-59}, {0x1.04525e0fc2fcbp+0, 0x1.12487a5507f68p-5, -0x1.804ad31b5f952p-61},
    {0x1.03ce4584b19a0p+0, 0x1.e38ce30333100p-6, -0x1.147b45033e1b4p-60}, {0x1.034ab2c50040dp+0, 0x1.a2a9c6c17044dp-6, -0x1.35b4d1c8470b4p-65},
    {0x1.02c7a505cffbfp+0, 0x1.61e77e8b53f9fp-6, 0x1.a2a0e2a1967efp-60}, {0x1.02451b7ddb2d2p+0, 0x1.2145e939ef1bcp-6, 0x1.47189d3ff66bfp-60},
    {0x1.01c315657186bp+0, 0x1.c189cbb0e283fp-7, 0x1.bb69dea7ecc2cp-61}, {0x1.014191f674111p+0, 0x1.40c8a7478788dp-7, -0x1.e20f8fffe770ap-61},
    {0x1.00c0906c513cfp+0, 0x1.809048289860ap-8, -0x1.6958f3f3b017bp-64}, {0x1.0000000000000p+0, 0x0.0p+0, 0x0.0p+0},
};

#ifndef CLODE_POLAR_HOST_CHECK
__shared__ ClodePolarEntry clode_polar_smem[256];
// every thread of the block, before any thread leaves the kernel (kernel prologue, with the exp table)
static __device__ __forceinline__ void clode_stage_polar_table()
{
    for (unsigned int j = threadIdx.x; j < 256u; j += blockDim.x)
        clode_polar_smem[j] = clode_polar_table[j];
    __syncthreads();
}
#define CLODE_POLAR_ENTRY(j) clode_polar_smem[j]
static __device__ __forceinline__ double clode_rsqrt_seed(const double u)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(u));
    return y;
}
#else
#define CLODE_POLAR_ENTRY(j) clode_polar_table[j]
#endif

__constant__ double clode_polar_c[8] = {
    -0x1.62e42fefa3800p+0,  // -2 ln2, high part (42 significant bits: times |e'| < 2^11 exactly)
    -0x1.ef35793c76730p-44, // -2 ln2, low part
    1.0 / 3.0, -2.0 / 5.0, 0.5, -2.0 / 3.0, // -2 log1p(r) = -2 r + r^2 (1 - 2/3 r + 1/2 r^2 - 2/5 r^3 + 1/3 r^4)
    0.375, 0.5};            // (1 - e)^(-1/2) = 1 + e/2 + 3/8 e^2 + ...

// sqrt(-2 log(q) / q) for 0 < q < 1 (what the polar method's loop hands over)
static __device__ __forceinline__ double clode_polar_scale(const double q)
{
    const double *c = clode_polar_c;
    const int hi = __double2hiint(q);
    const int e = ((hi + 0x00080000) >> 20) - 1023;                  // exponent after folding [1.5, 2) down to [0.75, 1)
    const double m = __hiloint2double(hi - (e << 20), __double2loint(q)); // q 2^-e in [0.75, 1.5)
    const ClodePolarEntry t = CLODE_POLAR_ENTRY((hi >> 12) & 255);
    const double r = fma(m, t.rc, -1.0);
    const double ed = (double)e;
    const double q4 = fma(r, fma(r, fma(r, fma(r, c[2], c[3]), c[4]), c[5]), 1.0);
    const double p = fma(r * r, q4, -2.0 * r);
    const double s = fma(ed, c[0], t.hi) + (p + fma(ed, c[1], t.lo)); // -2 log q  > 0
    const double u = s * q;
    const double y = clode_rsqrt_seed(u);
    const double d = fma(-(u * y), y, 1.0);
    return s * fma(y, d * fma(d, c[6], c[7]), y);
}
#endif // CLODE_FAST_POLAR_CUH
