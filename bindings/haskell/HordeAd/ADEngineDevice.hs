{-# LANGUAGE DataKinds, FlexibleContexts, GADTs, RankNTypes, ScopedTypeVariables, TypeApplications, TypeFamilies #-}
-- | Twins of @revInterpretArtifact@ / @vjpInterpretArtifact@ / @jvpInterpretArtifact@ (reference
-- ADEngine.hs:266-288, 420-470) that evaluate a gradient artifact on the B200 carrier, plus the device training-loop
-- plumbing (optimizers, gradient exchange, CUDA-graph capture) that has no counterpart above the tensor boundary.
--
-- The artifact itself -- produced ONCE by @revArtifactAdapt@ from the unchanged user program, simplified by
-- @simplifyArtifactRev@ -- is interpreted by the unchanged polymorphic @interpretAstPrimal@ with
-- @target ~ Device@.  With 'recordArtifact' the interpretation is RECORDED below the ABI instead of executed
-- (@hb_lazy_begin@ / @hb_lazy_end@): the library's contraction pass (csrc/lazy_rewrite.cu, the device sibling of
-- @contractAst@, AstTraverse.hs:290-699) rewrites the gather / sdot1In / kargMax program into fused conv / relu /
-- max-pool / loss nodes, and 'runRecorded' replays the result -- inside one CUDA graph when asked to.
--
-- NOT type-checked here (no GHC in the build image); horde-ad_b200/train.py is the executable image.
module HordeAd.ADEngineDevice
  ( -- * Evaluating artifacts on the device
    revInterpretArtifactDevice, vjpInterpretArtifactDevice, jvpInterpretArtifactDevice
    -- * Recording an artifact below the ABI (device contraction pass) and replaying it
  , Recorded, recordArtifact, runRecorded, freeRecorded, recordedStats
    -- * Whole-step capture
  , DeviceGraph, captureGraph, launchGraph, freeGraph
    -- * Optimizers and data-parallel exchange
  , sgdStepDevice, adamStepDevice, flattenBucket, allReduceAdamStep, allReduceSum
  , commInit, commUniqueId, commSetGlobalObjective
  ) where

import Prelude

import Control.Exception (onException)
import Data.ByteString (ByteString)
import Data.ByteString qualified as BS
import Data.ByteString.Unsafe qualified as BSU
import Foreign.C.Types (CInt (..))
import Foreign.ForeignPtr (ForeignPtr, newForeignPtr_, withForeignPtr)
import Foreign.Marshal.Alloc (alloca, allocaBytes)
import Foreign.Marshal.Array (withArrayLen)
import Foreign.Ptr (Ptr, castPtr, nullPtr)
import Foreign.Storable (peek)

import HordeAd.Core.Ast (AstArtifactFwd (..), AstArtifactRev (..))
import HordeAd.Core.AstEnv (emptyEnv, extendEnv)
import HordeAd.Core.AstInterpret (interpretAstPrimal)
import HordeAd.Core.DeviceFFI
import HordeAd.Core.OpsDevice
import HordeAd.Core.TensorKind
import HordeAd.Core.Types

-- | @revInterpretArtifact@ (ADEngine.hs:266-276) with a 'Device' environment: value and gradient of the objective
-- at @parameters@; @artDerivativeRev@ is evaluated with the incoming cotangent bound to @dt@ (default 1).
revInterpretArtifactDevice
  :: forall x z. AstArtifactRev x z -> Device x -> Maybe (Device (ADTensorKind z)) -> (Device (ADTensorKind x), Device z)
revInterpretArtifactDevice AstArtifactRev{..} parameters mdt =
  let azftk = adFTK (ftkAst artPrimalRev)
      env = extendEnv artVarDomainRev parameters emptyEnv
      dt = case mdt of
        Just a -> a
        Nothing -> treplTarget 1 azftk
      envDt = extendEnv artVarDtRev dt env
      gradient = interpretAstPrimal envDt artDerivativeRev
      primal = interpretAstPrimal env artPrimalRev
  in (gradient, primal)

vjpInterpretArtifactDevice
  :: forall x z. AstArtifactRev x z -> Device x -> Device (ADTensorKind z) -> (Device (ADTensorKind x), Device z)
vjpInterpretArtifactDevice art parameters dt = revInterpretArtifactDevice art parameters (Just dt)

-- | @jvpInterpretArtifact@ (ADEngine.hs:455-470).
jvpInterpretArtifactDevice
  :: forall x z. AstArtifactFwd x z -> Device x -> Device (ADTensorKind x) -> (Device (ADTensorKind z), Device z)
jvpInterpretArtifactDevice AstArtifactFwd{..} parameters ds =
  let env = extendEnv artVarDomainFwd parameters emptyEnv
      envD = extendEnv artVarDsFwd ds env
  in (interpretAstPrimal envD artDerivativeFwd, interpretAstPrimal env artPrimalFwd)

-- ---------------------------------------------------------------------------------------------------------------
-- Recording: every compute entry point called between hb_lazy_begin and hb_lazy_end returns a VIRTUAL handle and
-- appends a node to the library's DAG; nothing is launched.  hb_lazy_end runs the contraction pass (`passes` /= 0).

newtype Recorded = Recorded (ForeignPtr HbLazy)

-- | Interprets @step@ (typically @\\_ -> revInterpretArtifactDevice art params Nothing@ followed by 'flattenBucket'
-- and 'allReduceAdamStep') in recording mode; @outputs@ selects the arrays whose values the caller will read.
recordArtifact :: IO [DevArr] -> IO (Recorded, [DevArr])
recordArtifact step = do
  hbCheck c_hb_lazy_begin
  outs <- step `onException` c_hb_lazy_abort
  prog <- withMany' outs $ \ps -> withArrayLen ps $ \n pp -> alloca $ \pout -> do
    hbCheck (c_hb_lazy_end (fromIntegral n) pp 1 pout)
    peek pout
  fp <- newForeignPtr_ prog
  return (Recorded fp, outs)

runRecorded :: Recorded -> IO ()
runRecorded (Recorded fp) = withForeignPtr fp (hbCheck . c_hb_lazy_run)

freeRecorded :: Recorded -> IO ()
freeRecorded (Recorded fp) = withForeignPtr fp (hbCheck . c_hb_lazy_free)

-- | (nodes recorded, nodes left after the contraction pass): 55 -> 21 for MnistCnnShaped2's gradient step.
recordedStats :: Recorded -> IO (Int, Int)
recordedStats (Recorded fp) = withForeignPtr fp $ \p -> alloca $ \pr -> alloca $ \pl -> do
  hbCheck (c_hb_lazy_stats p pr pl)
  (,) <$> (fromIntegral <$> peek pr) <*> (fromIntegral <$> peek pl)

withMany' :: [DevArr] -> ([Ptr HbArray] -> IO a) -> IO a
withMany' [] k = k []
withMany' (a : as) k = withDev a $ \p -> withMany' as (k . (p :))

-- ---------------------------------------------------------------------------------------------------------------
-- CUDA-graph capture of a whole training step (SURVEY.md 8f-2): the launches issued by @body@ on the library stream
-- are captured instead of executed; addresses baked into the graph stay valid (private allocator pool).

newtype DeviceGraph = DeviceGraph (Ptr HbGraph)

captureGraph :: IO () -> IO DeviceGraph
captureGraph body = do
  hbCheck c_hb_graph_begin
  body `onException` alloca (\pg -> c_hb_graph_end pg)
  alloca $ \pg -> do
    hbCheck (c_hb_graph_end pg)
    DeviceGraph <$> peek pg

launchGraph :: DeviceGraph -> IO ()
launchGraph (DeviceGraph g) = hbCheck (c_hb_graph_launch g)

freeGraph :: DeviceGraph -> IO ()
freeGraph (DeviceGraph g) = hbCheck (c_hb_graph_free g)

-- ---------------------------------------------------------------------------------------------------------------
-- Device twins of OptimizerTools.hs (updateWithGradient :20-48, updateWithGradientAdam :134-232) on flat buffers,
-- and the gradient exchange of the batch-sharded training loop (SURVEY.md 8e).

-- | p := p - gamma * g, one fused pass.
sgdStepDevice :: Double -> DevArr -> DevArr -> IO ()
sgdStepDevice gamma p g = withDev p $ \pp -> withDev g $ \pg -> hbCheck (c_hb_sgd_step pp pg (realToFrac gamma))

-- | One fused pass over the seven streams (p, m, v read + written, g read); @betaOne ^ t@ as GHC's @(^)@ computes it.
adamStepDevice :: DevArr -> DevArr -> DevArr -> DevArr -> Int -> IO ()
adamStepDevice p m v g t = withDev p $ \pp -> withDev m $ \pm -> withDev v $ \pv -> withDev g $ \pg ->
  hbCheck (c_hb_adam_step pp pm pv pg 1e-3 0.9 0.999 1e-7 (fromIntegral t))

-- | The gradient leaves (in parameter order) and the loss in ONE flat bucket: one launch (@hb_flatten_concat@).
flattenBucket :: [DevArr] -> DevArr
flattenBucket leaves = wrapOut $ \out -> withMany' leaves $ \ps -> withArrayLen ps $ \n pp ->
  c_hb_flatten_concat (fromIntegral n) pp out

-- | The exchange AND the optimizer step in one kernel over NVLink-mapped peer buffers (csrc/comm.cu): every rank
-- pushes its bucket into a slot of every peer, the peers add the slots in rank order (bit-identical replicas),
-- scale, and apply Adam; the step counter lives on the device so the call replays from a CUDA graph.
allReduceAdamStep :: DevArr -> Int -> DevArr -> DevArr -> DevArr -> DevArr -> Double -> IO ()
allReduceAdamStep bucket nParams p m v tCounter scale =
  withDev bucket $ \pb -> withDev p $ \pp -> withDev m $ \pm -> withDev v $ \pv -> withDev tCounter $ \pt ->
    hbCheck (c_hb_allreduce_adam_step pb (fromIntegral nParams) pp pm pv pt 1e-3 0.9 0.999 1e-7 (realToFrac scale))

-- | NCCL all-reduce (sum) of a bucket followed by a scale: the comparison arm of 'allReduceAdamStep'.
allReduceSum :: DevArr -> Double -> IO ()
allReduceSum bucket scale = withDev bucket $ \pb -> hbCheck (c_hb_allreduce_sum pb (realToFrac scale))

-- | Rank 0 creates the id; the host program ships it to the other ranks (its own control plane), then every rank
-- calls 'commInit'.  One process (one GHC runtime) per GPU.
commUniqueId :: IO ByteString
commUniqueId = allocaBytes 128 $ \p -> do
  hbCheck (c_hb_comm_unique_id p)
  BS.packCStringLen (castPtr p, 128)

commInit :: Int -> Int -> ByteString -> IO ()
commInit rank world uid
  | world == 1 = hbCheck (c_hb_comm_init 0 1 nullPtr)
  | otherwise = BSU.unsafeUseAsCString uid $ \p -> hbCheck (c_hb_comm_init (fromIntegral rank) (fromIntegral world) (castPtr p))

-- | Global-objective mode: @minimum u@ and @sum expU@ of @lossSoftMaxCrossEntropyS@ (CommonShapedOps.hs:89-111)
-- range over all shards, so an N-rank step equals the reference's one-device step on the whole batch.
commSetGlobalObjective :: Bool -> IO ()
commSetGlobalObjective on = hbCheck (c_hb_comm_set_global_objective (if on then 1 else 0))
