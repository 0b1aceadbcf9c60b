V="python tools/verify_sign_only.py --region 512"
$V --seeds 7 --origin 100000,0,-77777
$V --seeds 7 --origin 3000000,2,3000000
$V --seeds 3 --origin=-20000000,0,15000000
$V --seeds 5 --freq 0.05 --lacunarity 0.6 --gain -2.0 --origin 0,3,0
$V --seeds 5 --freq 0.1 --lacunarity 1.0 --gain 1.0
$V --seeds 9 --freq 0.02 --lacunarity 2.0 --gain 0.5 --origin=-300,0,800
$V --seeds 1 --freq -0.09 --lacunarity -0.5 --gain 2.5
$V --seeds 1 --freq 0.3 --origin 0,-1,0
$V --seeds 2 --freq 0.004 --lacunarity 0.5 --gain 3.5
