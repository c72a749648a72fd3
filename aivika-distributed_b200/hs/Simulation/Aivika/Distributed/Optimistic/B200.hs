{-# LANGUAGE ForeignFunctionInterface #-}
{-# LANGUAGE RecordWildCards #-}

-- |
-- Module     : Simulation.Aivika.Distributed.Optimistic.B200
-- Stability  : experimental (NOT COMPILED HERE: this image has no GHC; see INTEGRATION.md)
--
-- Haskell side of the B200 Time Warp backend: a thin FFI layer over @include/twb200.h@ that keeps the
-- reference's surface for the batched path -- 'runDIOBatch' (shape of 'runDIO',
-- Internal/DIO.hs:327-328), 'DIOParams' (Internal/DIO.hs:70-110: 'dioTimeHorizon' is forwarded),
-- 'enqueueMessageB200' (Message.hs:48-69) and the Time Server role ('globalTimeB200' = the value a
-- @GlobalTimeMessage@ would carry, Internal/TimeServer.hs:384-403).
--
-- Arbitrary Haskell closures cannot run on the device, so a batch of homogeneous LPs is described by a
-- compiled model ('ModelPHOLD' / 'ModelMachRep2' / 'ModelMachRep1') instead of a @Simulation DIO a@ value.
-- Under an external Time Server (Internal/TimeServer.hs) a batch behaves like one logical process:
-- 'attachTimeServerB200', 'localTimeB200' (the @LocalTimeMessage@ payload) and 'globalTimeMessageB200'.
-- All calls are @safe@ so the RTS keeps running other Haskell threads during 'twRun'.
module Simulation.Aivika.Distributed.Optimistic.B200
       (B200Model(..), B200Specs(..), B200Params(..), defaultB200Params, RunStats(..), MachRepResult(..),
        B200Engine, runDIOBatch, newB200Engine, freeB200Engine, loadInitialMessages,
        enqueueMessageB200, runEventInStopTimeB200, globalTimeB200, readMachRep2,
        attachTimeServerB200, localTimeB200, globalTimeMessageB200) where

import Control.Exception (throwIO, ErrorCall(..), bracket)
import Control.Monad (when)
import Data.Int
import Data.Word
import Foreign
import Foreign.C.String
import Foreign.C.Types

-- | Opaque @tw_engine*@.
newtype B200Engine = B200Engine (Ptr ())

data B200Model
  = ModelPHOLD { pholdLPs :: Word32, pholdPopulation :: Word32, pholdRemotePct :: Word32
               , pholdLookahead :: Double, pholdMean :: Double, pholdSeed :: Word64
               , pholdHot :: Word32, pholdHotPct :: Word32 }
  | ModelMachRep2 { mrReplicas :: Word32, mrMeanUpTime :: Double, mrMeanRepairTime :: Double
                  , mrDelta :: Double, mrSeed1 :: Int64, mrSeed2 :: Int64, mrSeedStride :: Int64 }
  | ModelMachRep1 { m1Replicas :: Word32, m1MeanUpTime :: Double, m1MeanRepairTime :: Double
                  , m1Seed1 :: Int64, m1Seed2 :: Int64, m1SeedStride :: Int64 }     -- tests/MachRep1.hs

data B200Specs = B200Specs { spcStartTime :: Double, spcStopTime :: Double }

-- | The hot-path fields of 'DIOParams' plus the engine knobs.
data B200Params = B200Params
  { dioTimeHorizon :: Maybe Double      -- ^ 'Nothing' = unbounded optimism (the reference default)
  , queueCapacity :: Word32, maxEventsPerStep :: Word32, stepsPerLaunch :: Word32 }

defaultB200Params :: B200Params
defaultB200Params = B200Params Nothing 128 4 16

data RunStats = RunStats { rsGVT :: Double, rsDeviceMs :: Double, rsSteps :: Word64, rsCommitted :: Word64
                         , rsExecuted :: Word64, rsRolledBack :: Word64, rsRollbacks :: Word64 } deriving Show

data MachRepResult = MachRepResult { totalUpTime :: Double, nRep :: Int64, nImmedRep :: Int64
                                   , upTimeProp :: Double, immedProp :: Double } deriving Show

-- sizes / offsets of the C structs (include/twb200.h; checked by tests/test_abi.py against gcc)
sizeofParams, sizeofEvent, sizeofRunStats, sizeofMachRep :: Int
sizeofParams = 296; sizeofEvent = 48; sizeofRunStats = 120; sizeofMachRep = 40

foreign import ccall safe "tw_create"  c_tw_create  :: Ptr () -> Ptr (Ptr ()) -> IO CInt
foreign import ccall safe "tw_destroy" c_tw_destroy :: Ptr () -> IO CInt
foreign import ccall safe "tw_reset"   c_tw_reset   :: Ptr () -> IO CInt
foreign import ccall safe "tw_enqueue_message"  c_tw_enqueue_message  :: Ptr () -> Ptr () -> IO CInt
foreign import ccall safe "tw_enqueue_messages" c_tw_enqueue_messages :: Ptr () -> Ptr () -> CSize -> IO CInt
-- | runEventInStartTime of every LP on the device (the model's own start-time events; no host buffer).
foreign import ccall safe "tw_populate" c_tw_populate :: Ptr () -> IO CInt
-- | Largest ring fill of an LP with a long pending region (the fixed-capacity form of the isEventOverflow thresholds).
foreign import ccall safe "tw_max_ring_occupancy" c_tw_max_ring_occupancy :: Ptr () -> Ptr Word32 -> IO CInt
foreign import ccall safe "tw_run"     c_tw_run     :: Ptr () -> CDouble -> Ptr () -> IO CInt
foreign import ccall safe "tw_gvt"     c_tw_gvt     :: Ptr () -> Ptr CDouble -> IO CInt
foreign import ccall safe "tw_read_machrep" c_tw_read_machrep :: Ptr () -> Ptr () -> CSize -> IO CInt
foreign import ccall safe "tw_model_initial_events" c_tw_model_initial_events :: Ptr () -> Ptr () -> CSize -> Ptr CSize -> IO CInt
foreign import ccall safe "tw_attach_time_server" c_tw_attach_time_server :: Ptr () -> CInt -> IO CInt
foreign import ccall safe "tw_local_time" c_tw_local_time :: Ptr () -> Ptr CDouble -> IO CInt
foreign import ccall safe "tw_global_time_message" c_tw_global_time_message :: Ptr () -> CDouble -> Ptr CInt -> IO CInt
foreign import ccall unsafe "tw_last_error" c_tw_last_error :: Ptr () -> IO CString

-- | Non-zero status becomes an 'error', like the invariants of the reference (e.g. Internal/Event.hs:374).
check :: Ptr () -> CInt -> IO ()
check h rc = when (rc /= 0) $ do
  msg <- c_tw_last_error h >>= peekCString
  throwIO (ErrorCall ("twb200: " ++ msg))

-- | Fill a zeroed @tw_params@.
pokeParams :: Ptr () -> B200Model -> B200Params -> B200Specs -> IO ()
pokeParams p model B200Params{..} B200Specs{..} = do
  fillBytes p 0 sizeofParams
  pokeByteOff p 8  (0 :: Word32)            -- rank
  pokeByteOff p 12 (1 :: Word32)            -- world
  pokeByteOff p 16 (-1 :: Int32)            -- device
  pokeByteOff p 20 queueCapacity
  pokeByteOff p 24 maxEventsPerStep
  pokeByteOff p 28 stepsPerLaunch
  pokeByteOff p 32 spcStartTime
  pokeByteOff p 40 spcStopTime
  pokeByteOff p 48 (maybe (-1) id dioTimeHorizon :: Double)
  pokeByteOff p 68 (1 :: Word32)            -- flags: TW_FLAG_SKIP_OUTDATED (dioAllowSkippingOutdatedMessage = True, Internal/DIO.hs:206)
  case model of
    ModelPHOLD{..} -> do
      pokeByteOff p 0 (1 :: Word32); pokeByteOff p 4 pholdLPs
      pokeByteOff p 72 pholdPopulation; pokeByteOff p 76 pholdRemotePct
      pokeByteOff p 80 pholdHot; pokeByteOff p 84 pholdHotPct
      pokeByteOff p 88 pholdLookahead; pokeByteOff p 96 pholdMean; pokeByteOff p 104 pholdSeed
    ModelMachRep2{..} -> do
      pokeByteOff p 0 (2 :: Word32); pokeByteOff p 4 (3 * mrReplicas)
      pokeByteOff p 112 mrMeanUpTime; pokeByteOff p 120 mrMeanRepairTime; pokeByteOff p 128 mrDelta
      pokeByteOff p 136 mrSeed1; pokeByteOff p 144 mrSeed2; pokeByteOff p 152 mrSeedStride
    ModelMachRep1{..} -> do
      pokeByteOff p 0 (3 :: Word32); pokeByteOff p 4 (3 * m1Replicas)
      pokeByteOff p 112 m1MeanUpTime; pokeByteOff p 120 m1MeanRepairTime; pokeByteOff p 128 (0 :: Double)
      pokeByteOff p 136 m1Seed1; pokeByteOff p 144 m1Seed2; pokeByteOff p 152 m1SeedStride

-- | @newEventQueue@ for a whole batch of LPs (Internal/Event.hs:92-116).
newB200Engine :: B200Model -> B200Params -> B200Specs -> IO B200Engine
newB200Engine model ps specs =
  allocaBytes sizeofParams $ \p -> alloca $ \out -> do
    pokeParams p model ps specs
    rc <- c_tw_create p out
    check nullPtr rc
    B200Engine <$> peek out

freeB200Engine :: B200Engine -> IO ()
freeB200Engine (B200Engine h) = c_tw_destroy h >>= check h

-- | The model's start-time messages (@runEventInStartTime machine@ / the PHOLD population), as one bulk message.
loadInitialMessages :: B200Engine -> B200Model -> B200Params -> B200Specs -> IO ()
loadInitialMessages (B200Engine h) model ps specs =
  allocaBytes sizeofParams $ \p -> do
    pokeParams p model ps specs
    let cap = case model of
                ModelPHOLD{..}    -> fromIntegral pholdLPs * fromIntegral pholdPopulation
                ModelMachRep2{..} -> 2 * fromIntegral mrReplicas
                ModelMachRep1{..} -> 2 * fromIntegral m1Replicas
    allocaBytes (cap * sizeofEvent) $ \evs -> alloca $ \pn -> do
      c_tw_model_initial_events p evs (fromIntegral cap) pn >>= check h
      n <- peek pn
      c_tw_enqueue_messages h evs n >>= check h

-- | @enqueueMessage pid t a@ (Message.hs:48-69) issued by the host at the start time.
enqueueMessageB200 :: B200Engine -> Word32 -> Word32 -> Double -> Double -> Word32 -> Double -> IO ()
enqueueMessageB200 (B200Engine h) src dst sendTime recvTime kind value =
  allocaBytes sizeofEvent $ \e -> do
    fillBytes e 0 sizeofEvent
    pokeByteOff e 0 recvTime; pokeByteOff e 8 sendTime; pokeByteOff e 24 src; pokeByteOff e 28 dst
    pokeByteOff e 32 kind; pokeByteOff e 40 value
    c_tw_enqueue_message h e >>= check h

-- | @runEventInStopTime@ (Internal/Event.hs:124-129, 743-752): returns when GVT has passed the stop time.
runEventInStopTimeB200 :: B200Engine -> Double -> IO RunStats
runEventInStopTimeB200 (B200Engine h) t =
  allocaBytes sizeofRunStats $ \st -> do
    c_tw_run h (realToFrac t) st >>= check h
    RunStats <$> peekByteOff st 0 <*> peekByteOff st 8 <*> peekByteOff st 16 <*> peekByteOff st 24
             <*> peekByteOff st 32 <*> peekByteOff st 40 <*> peekByteOff st 48

-- | The Time Server's answer (@GlobalTimeMessage@).
globalTimeB200 :: B200Engine -> IO Double
globalTimeB200 (B200Engine h) = alloca $ \g -> do
  c_tw_gvt h g >>= check h
  realToFrac <$> peek g

-- | @registerDIO@ with an external Time Server (Internal/DIO.hs:267-283): GVT follows its @GlobalTimeMessage@s from now on.
attachTimeServerB200 :: B200Engine -> IO ()
attachTimeServerB200 (B200Engine h) = c_tw_attach_time_server h 1 >>= check h

-- | The payload of the @LocalTimeMessage@ sent in reply to @ComputeLocalTimeMessage@ (Internal/Event.hs:425-436).
localTimeB200 :: B200Engine -> IO Double
localTimeB200 (B200Engine h) = alloca $ \t -> do
  c_tw_local_time h t >>= check h
  realToFrac <$> peek t

-- | @GlobalTimeMessage t@ -> @updateGlobalTime@ (Internal/Event.hs:437-449, 528-540); 'False' = ignored (above the local time).
globalTimeMessageB200 :: B200Engine -> Double -> IO Bool
globalTimeMessageB200 (B200Engine h) t = alloca $ \ok -> do
  c_tw_global_time_message h (realToFrac t) ok >>= check h
  (/= 0) <$> peek ok

-- | The pair returned by @masterModel@ (tests/MachRep2DistributedReproducible.hs:214-229), per replica.
readMachRep2 :: B200Engine -> Int -> IO [MachRepResult]
readMachRep2 (B200Engine h) n =
  allocaBytes (n * sizeofMachRep) $ \buf -> do
    c_tw_read_machrep h buf (fromIntegral n) >>= check h
    mapM (\i -> let o = i * sizeofMachRep in
                MachRepResult <$> peekByteOff buf o <*> peekByteOff buf (o + 8) <*> peekByteOff buf (o + 16)
                              <*> peekByteOff buf (o + 24) <*> peekByteOff buf (o + 32)) [0 .. n - 1]

-- | Shape of @runDIO m ps timeServerId@: build the batch, load the start-time messages, run to the stop time.
runDIOBatch :: B200Model -> B200Params -> B200Specs -> IO (RunStats, B200Engine)
runDIOBatch model ps specs = do
  e <- newB200Engine model ps specs
  loadInitialMessages e model ps specs
  st <- runEventInStopTimeB200 e (spcStopTime specs)
  return (st, e)
