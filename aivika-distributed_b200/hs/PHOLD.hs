{-# LANGUAGE DeriveGeneric #-}
{-# LANGUAGE DeriveDataTypeable #-}

-- PHOLD on the REFERENCE engine (aivika-distributed), NOT COMPILED IN THIS REPOSITORY'S IMAGE (no GHC).
--
-- Purpose (SURVEY.md 8d, DESIGN.md 6): anyone with GHC and the reference's dependencies can
--   (a) produce the true reference number for BASELINE configs[2..4] (committed events / second of the Haskell
--       engine, -O2 -threaded, +RTS -N), and
--   (b) pin this repository's oracle to the real engine: the program prints the total number of executed events
--       and the wrapping sum of the per-LP trace hashes, which must equal `committed` / `hash_sum` of
--       oracle_phold_run (oracle/oracle.c) and `tw_run_stats.committed` / `.trace_hash_sum` of libtwb200.so for
--       the same parameters.
--
-- The model is the one of SURVEY.md App. C / DESIGN.md 7, written against the reference's public API exactly like
-- tests/MachRep1.hs and tests/MachRep2DistributedReproducible.hs: every LP is `runDIO` + `runSimulation`, receives
-- jobs through `messageReceived`, keeps its counters in undoable `Ref DIO`s and sends one job per executed job with
-- `enqueueMessage pid (t + delay)`.  All arithmetic is the bit-exact definition of csrc/tw_math.cuh /
-- oracle/oracle_math.h: splitmix64 finaliser, counter-based draws, u01 = ((d >> 12) + 1/2) * 2^-52, the fixed-order
-- software logarithm (fdlibm coefficients, no FMA), `-log u * mean`.
--
-- Build:  ghc -O2 -threaded PHOLD.hs
-- Run:    ./PHOLD <n_lps> <stop_time> <remote_pct> [<time_horizon> | none] +RTS -N
--         (population 16, lookahead 0.1, mean 1.0, seed 1 = bench.py's c3/c4; all LPs on the local node)
-- Output: committed=<n> hash_sum=0x<hex> seconds=<wall> events_per_sec=<rate>

import System.Environment (getArgs)
import System.IO
import Data.Array
import Data.Bits
import Data.IORef
import Data.Word
import Data.Typeable
import Data.Binary
import Data.Time.Clock
import Numeric (showHex)
import GHC.Generics
import GHC.Float (castDoubleToWord64, castWord64ToDouble)

import Control.Monad
import Control.Monad.Trans
import Control.Concurrent.MVar
import System.IO.Unsafe (unsafePerformIO)
import qualified Control.Distributed.Process as DP
import Control.Distributed.Process.Node (initRemoteTable)
import Control.Distributed.Process.Backend.SimpleLocalnet

import Simulation.Aivika.Trans
import Simulation.Aivika.Distributed

-- ---------------------------------------------------------------- bit-exact arithmetic (csrc/tw_math.cuh)

kGolden, kMulA, kMulB, kCtrMul :: Word64
kGolden = 0x9E3779B97F4A7C15
kMulA   = 0xBF58476D1CE4E5B9
kMulB   = 0x94D049BB133111EB
kCtrMul = 0xD1B54A32D192ED03

mixFin :: Word64 -> Word64
mixFin z0 = z2 `xor` (z2 `shiftR` 31)
  where z1 = (z0 `xor` (z0 `shiftR` 30)) * kMulA
        z2 = (z1 `xor` (z1 `shiftR` 27)) * kMulB

lpStreamKey :: Word64 -> Int -> Word64
lpStreamKey seed lp = mixFin (seed `xor` ((fromIntegral lp + 1) * kGolden))

streamDraw :: Word64 -> Word64 -> Word64
streamDraw key ctr = mixFin (key + ctr * kCtrMul)

mulHi64 :: Word64 -> Word64 -> Word64
mulHi64 a b = fromIntegral ((toInteger a * toInteger b) `shiftR` 64)

-- | uniform on (0,1): 52 bits + 1/2, exact
uniform01 :: Word64 -> Double
uniform01 d = (fromIntegral (d `shiftR` 12) + 0.5) * 2.220446049250313e-16      -- 2^-52

-- | natural logarithm of a positive normal double, fixed operation order (tw_math.cuh log_exact)
logExact :: Double -> Double
logExact x
  | b0 == 0   = -1 / 0
  | otherwise = dk * ln2hi - ((hfsq - inner) - f)
  where
    ln2hi = 6.93147180369123816490e-01; ln2lo = 1.90821492927058770002e-10
    c1 = 6.666666666666735130e-01; c2 = 3.999999999940941908e-01; c3 = 2.857142874366239149e-01
    c4 = 2.222219843214978396e-01; c5 = 1.818357216161805012e-01; c6 = 1.531383769920937332e-01
    c7 = 1.479819860511658591e-01
    b0 = castDoubleToWord64 x
    k0 = fromIntegral (b0 `shiftR` 52) - 1023 :: Int
    m  = b0 .&. 0x000FFFFFFFFFFFFF
    (k, b) | m > 0x6A09E667F3BCC = (k0 + 1, m .|. 0x3FE0000000000000)
           | otherwise           = (k0,     m .|. 0x3FF0000000000000)
    f  = castWord64ToDouble b - 1.0
    dk = fromIntegral k :: Double
    s  = f / (2.0 + f)
    z  = s * s
    w  = z * z
    t1 = w * (c2 + w * (c4 + w * c6))
    t2 = z * (c1 + w * (c3 + w * (c5 + w * c7)))
    r  = t2 + t1
    hfsq  = (0.5 * f) * f
    inner = s * (hfsq + r) + dk * ln2lo

expoVariate :: Double -> Double -> Double
expoVariate u mu = negate (logExact u * mu)

traceFold :: Word64 -> Double -> Int -> Word64 -> Word64
traceFold h recvTs src ctr = mixFin ((h `xor` x) + kGolden)
  where x = castDoubleToWord64 recvTs `xor` (ctr * kMulB) `xor` (fromIntegral src * kMulA)

-- ---------------------------------------------------------------- the model

data PholdParams = PholdParams { pN :: Int, pPopulation :: Int, pRemotePct :: Word64, pLookahead :: Double
                               , pMean :: Double, pSeed :: Word64, pStop :: Double }

-- | The payload: sender LP and the sender's send counter (the message identity the trace hash folds).
data Job = Job Int Word64 deriving (Eq, Ord, Show, Typeable, Generic)
instance Binary Job

specsFor :: PholdParams -> Specs DIO
specsFor p = Specs { spcStartTime = 0.0, spcStopTime = pStop p, spcDT = 1.0,
                     spcMethod = RungeKutta4, spcGeneratorType = SimpleGenerator }

-- | LP @i@: returns (executed events, trace hash) at the stop time.
pholdModel :: PholdParams -> MVar (Array Int DP.ProcessId) -> Int -> Simulation DIO (Word64, Word64)
pholdModel p tableVar i =
  do let key = lpStreamKey (pSeed p) i
         pop = fromIntegral (pPopulation p) :: Word64
     rngCtr  <- newRef pop
     sendCtr <- newRef pop
     nExec   <- newRef (0 :: Word64)
     hashRef <- newRef (mixFin (key + 1))
     -- the inbox ids of all LPs: filled in by `master` before any LP starts; read on first use (DIO offers no liftIO
     -- in Simulation, and the table is immutable once written, so the lazy read is safe)
     let table = unsafePerformIO (readMVar tableVar)
         n = fromIntegral (pN p) :: Word64

     runEventInStartTime $
       handleSignal messageReceived $ \(Job src ctr) ->
       do t  <- liftDynamics time
          modifyRef hashRef (\h -> traceFold h t src ctr)
          modifyRef nExec (+ 1)
          rc <- readRef rngCtr
          sc <- readRef sendCtr
          let d1 = streamDraw key rc
              d2 = streamDraw key (rc + 1)
              d3 = streamDraw key (rc + 2)
              delay  = pLookahead p + expoVariate (uniform01 d1) (pMean p)
              remote = n > 1 && (((d2 `shiftR` 32) * 100) `shiftR` 32) < pRemotePct p
              dst | remote    = let k = mulHi64 d3 (n - 1) in fromIntegral (k + (if k >= fromIntegral i then 1 else 0))
                  | otherwise = i
          writeRef rngCtr (rc + 4)
          writeRef sendCtr (sc + 1)
          enqueueMessage (table ! dst) (t + delay) (Job i sc)

     -- the initial population: P jobs to itself, send time 0
     runEventInStartTime $
       forM_ [0 .. pPopulation p - 1] $ \j ->
       enqueueMessage (table ! i) (expoVariate (uniform01 (streamDraw key (fromIntegral j))) (pMean p))
                      (Job i (fromIntegral j))

     runEventInStopTime $
       do a <- readRef nExec
          h <- readRef hashRef
          return (a, h)

-- ---------------------------------------------------------------- running N logical processes on the local node

master :: PholdParams -> Maybe Double -> Backend -> [DP.NodeId] -> DP.Process ()
master p horizon _backend _nodes =
  do let n  = pN p
         ps = defaultDIOParams { dioLoggingPriority = WARNING, dioTimeHorizon = horizon }
     timeServerId <- DP.spawnLocal $ timeServer n defaultTimeServerParams { tsLoggingPriority = WARNING }
     tableVar <- liftIO newEmptyMVar
     results  <- liftIO $ newIORef ([] :: [MVar (Word64, Word64)])
     -- runDIO hands out the inbox id before the model runs, so the id table can be completed first
     lps <- forM [0 .. n - 1] $ \i ->
       do let m = do registerDIO
                     a <- runSimulation (pholdModel p tableVar i) (specsFor p)
                     if i == 0 then terminateDIO else unregisterDIO
                     return a
          runDIO m ps timeServerId
     liftIO $ putMVar tableVar (listArray (0, n - 1) (map fst lps))
     t0 <- liftIO getCurrentTime
     forM_ (tail lps) $ \(_, proc) ->
       do v <- liftIO newEmptyMVar
          liftIO $ modifyIORef results (v :)
          DP.spawnLocal (proc >>= liftIO . putMVar v)
     r0 <- snd (head lps)
     rs <- liftIO $ readIORef results >>= mapM takeMVar
     t1 <- liftIO getCurrentTime
     let committed = sum (map fst (r0 : rs))
         hashSum   = sum (map snd (r0 : rs))                  -- wraps modulo 2^64
         secs      = realToFrac (diffUTCTime t1 t0) :: Double
     liftIO $ putStrLn $ "committed=" ++ show committed ++ " hash_sum=0x" ++ showHex hashSum ""
                       ++ " seconds=" ++ show secs ++ " events_per_sec=" ++ show (fromIntegral committed / secs)
     liftIO $ hFlush stdout

main :: IO ()
main = do
  args <- getArgs
  let (n, stop, rp, hz) = case args of
        [a, b, c]    -> (read a, read b, read c, Nothing)
        [a, b, c, d] -> (read a, read b, read c, if d == "none" then Nothing else Just (read d))
        _            -> (64, 5.0, 10, Nothing)
      p = PholdParams { pN = n, pPopulation = 16, pRemotePct = rp, pLookahead = 0.1, pMean = 1.0, pSeed = 1, pStop = stop }
  backend <- initializeBackend "localhost" "8088" initRemoteTable
  startMaster backend (master p hz backend)
