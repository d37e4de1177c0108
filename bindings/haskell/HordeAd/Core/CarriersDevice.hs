{-# LANGUAGE AllowAmbiguousTypes, DataKinds, FlexibleInstances, GADTs, KindSignatures, LambdaCase,
             MultiParamTypeClasses, ScopedTypeVariables, TypeApplications, TypeFamilies, UndecidableInstances #-}
-- | The device carrier: sibling of "HordeAd.Core.CarriersConcrete" (reference CarriersConcrete.hs:150-155,293)
-- for tensors that live in B200 HBM behind the C ABI of libhordeb200 (include/hordeb200.h).
--
-- Where 'Concrete' wraps ox-arrays values ('RepConcrete'), 'Device' wraps reference-counted @hb_array@ handles:
-- strided views over device storage (stride 0 = @sreplicate@, negative = @sreverse@), so the metadata-only
-- operations of ox-arrays stay metadata-only.  Scalars ('TKScalar') have three forms (SURVEY.md 7.2-2 option A and
-- 7.2-6): a host literal, a 0-d device array (results of @ssum0@, @sdot0@ ... that are never forced), or -- only
-- while an index function of @tsgather@ / @tsscatter@ is being reified -- a symbolic expression over the loop
-- variables, which "HordeAd.Core.OpsDevice" compiles to the library's index bytecode (@hb_ixprog_build@).
--
-- NOT type-checked in this repository (no GHC in the build image); the executable image of the same calls is
-- horde-ad_b200/device.py, which the test-suite drives through the same ABI.
module HordeAd.Core.CarriersDevice
  ( -- * Handles
    DevArr(..), withDev, wrapOut, wrapOut2, hbCheck, hbError
    -- * The carrier
  , Device(..), DevRep(..), HostScalar(..), DevBool(..)
    -- * Symbolic scalars (index-function reification)
  , IxExpr(..), FxExpr(..), BExpr(..), IxBin(..), IxUn(..), IxCmp(..)
    -- * Element types
  , HbElt(..), dtypeCode
    -- * Scalars
  , devScalarHost, devScalarToHost, devScalarAsArr, hostInt
    -- * Runtime
  , deviceInit, deviceSync
  ) where

import Control.Monad (when)
import Data.Int (Int16, Int32, Int64, Int8)
import Foreign.C.String (peekCString)
import Foreign.C.Types (CInt (..))
import Foreign.ForeignPtr (ForeignPtr, newForeignPtr, withForeignPtr)
import Foreign.Marshal.Alloc (alloca, allocaBytes)
import Foreign.Marshal.Utils (with)
import Foreign.Ptr (Ptr, castPtr, nullPtr)
import Foreign.Storable (Storable, peek)
import System.IO.Unsafe (unsafePerformIO)

import HordeAd.Core.DeviceFFI
import HordeAd.Core.TensorKind (TK (..))

-- | A device array handle.  The finaliser is @hb_release@, callable from any OS thread (GHC runs finalisers on
-- arbitrary capabilities; INTEGRATION.md section 5).
newtype DevArr = DevArr (ForeignPtr HbArray)

withDev :: DevArr -> (Ptr HbArray -> IO a) -> IO a
withDev (DevArr fp) = withForeignPtr fp

hbError :: IO a
hbError = allocaBytes 1024 $ \buf -> do
  _ <- c_hb_last_error buf 1024
  msg <- peekCString buf
  error ("libhordeb200: " ++ msg)   -- programmer errors only, as OpsConcrete.hs:129 / ADEngine.hs:194-196

hbCheck :: IO CInt -> IO ()
hbCheck act = do
  st <- act
  when (st /= 0) hbError

-- | Every producing entry point returns its result through an out-parameter.  Pure, like every 'Concrete' method:
-- the library is a deterministic function of its operands (fixed reduction orders, DESIGN.md section 3).
wrapOut :: (Ptr (Ptr HbArray) -> IO CInt) -> DevArr
wrapOut k = unsafePerformIO $ alloca $ \out -> do
  hbCheck (k out)
  DevArr <$> (newForeignPtr p_hb_release =<< peek out)
{-# NOINLINE wrapOut #-}

wrapOut2 :: (Ptr (Ptr HbArray) -> Ptr (Ptr HbArray) -> IO CInt) -> (DevArr, DevArr)
wrapOut2 k = unsafePerformIO $ alloca $ \o1 -> alloca $ \o2 -> do
  hbCheck (k o1 o2)
  a <- DevArr <$> (newForeignPtr p_hb_release =<< peek o1)
  b <- DevArr <$> (newForeignPtr p_hb_release =<< peek o2)
  return (a, b)
{-# NOINLINE wrapOut2 #-}

-- | Element types of the boundary (the seven dtypes of @hb_dtype@; 'GoodScalar' instances of the reference).
class (Storable r, Show r) => HbElt r where
  hbDtype :: CInt
  toHostScalar :: r -> HostScalar
  fromHostScalar :: HostScalar -> r
instance HbElt Double where { hbDtype = 0; toHostScalar = HDouble; fromHostScalar = \case { HDouble x -> x; h -> realToFrac (hostToDouble h) } }
instance HbElt Float  where { hbDtype = 1; toHostScalar = HFloat;  fromHostScalar = \case { HFloat x -> x;  h -> realToFrac (hostToDouble h) } }
instance HbElt Int64  where { hbDtype = 2; toHostScalar = HInt . fromIntegral; fromHostScalar = fromIntegral . hostToInt }
instance HbElt Int    where { hbDtype = 2; toHostScalar = HInt . fromIntegral; fromHostScalar = fromIntegral . hostToInt }
instance HbElt Int32  where { hbDtype = 3; toHostScalar = HInt . fromIntegral; fromHostScalar = fromIntegral . hostToInt }
instance HbElt Int16  where { hbDtype = 4; toHostScalar = HInt . fromIntegral; fromHostScalar = fromIntegral . hostToInt }
instance HbElt Int8   where { hbDtype = 5; toHostScalar = HInt . fromIntegral; fromHostScalar = fromIntegral . hostToInt }
instance HbElt Bool   where { hbDtype = 6; toHostScalar = HBool; fromHostScalar = \case { HBool b -> b; h -> hostToInt h /= 0 } }

dtypeCode :: forall r. HbElt r => CInt
dtypeCode = hbDtype @r

data HostScalar = HDouble !Double | HFloat !Float | HInt !Int64 | HBool !Bool
  deriving (Eq, Show)

hostToDouble :: HostScalar -> Double
hostToDouble = \case { HDouble x -> x; HFloat x -> realToFrac x; HInt n -> fromIntegral n; HBool b -> if b then 1 else 0 }
hostToInt :: HostScalar -> Int64
hostToInt = \case { HDouble x -> floor x; HFloat x -> floor x; HInt n -> n; HBool b -> if b then 1 else 0 }
hostInt :: Int -> DevRep
hostInt = DHost . HInt . fromIntegral

-- | Symbolic integer / real expressions: exactly the grammar of SURVEY.md appendix B that the reference's
-- vectoriser leaves inside gather / scatter index functions (Ast.hs:487-528,575-579,654-665).
data IxBin = IxAdd | IxSub | IxMul | IxQuot | IxRem | IxMin | IxMax deriving (Eq, Show)
data IxUn = IxNeg | IxAbs | IxSignum deriving (Eq, Show)
data IxCmp = CmpLeq | CmpLt | CmpEq deriving (Eq, Show)
data IxExpr
  = IxVar !Int                        -- ^ k-th component of the @shm@ index
  | IxConst !Int64
  | IxBinOp !IxBin IxExpr IxExpr
  | IxUnOp !IxUn IxExpr
  | IxIf BExpr IxExpr IxExpr          -- ^ @ifH@
  | IxTbl DevArr [IxExpr]             -- ^ @sindex0@ of an integer tensor at symbolic indices (OOB -> 0)
  | IxArgMax DevArr [IxExpr]          -- ^ @kargMax@ of the last axis at a symbolic prefix
  | IxArgMin DevArr [IxExpr]
  | IxFloor FxExpr                    -- ^ @kfloor@ of a real expression
data FxExpr
  = FxConst !Double
  | FxTbl DevArr [IxExpr]             -- ^ @sindex0@ of a real tensor
  | FxFromInt IxExpr                  -- ^ @kfromIntegral@
  | FxBin !Int FxExpr FxExpr          -- ^ an @hb_vmop@ real binary op code
  | FxUn !Int FxExpr                  -- ^ an @hb_unop@ code
data BExpr
  = BConst !Bool
  | BCmp !IxCmp IxExpr IxExpr
  | BCmpF !IxCmp FxExpr FxExpr
  | BAnd BExpr BExpr | BOr BExpr BExpr | BNot BExpr

-- | Runtime representation; the type index of 'Device' is a phantom, as the handles carry shape and dtype.
data DevRep
  = DArr !DevArr                      -- ^ any ranked / shaped / mixed tensor, or a 0-d scalar left on the device
  | DHost !HostScalar                 -- ^ a scalar known on the host
  | DSymI IxExpr                      -- ^ symbolic integer scalar (only inside index-function reification)
  | DSymF FxExpr                      -- ^ symbolic real scalar
  | DSymT !DevArr [IxExpr]            -- ^ a tensor indexed by a symbolic prefix (@tsindex t ix@ inside an index
                                      --   function): only @tsindex0@ / @tkargMax@ / @tkargMin@ may consume it
  | DPair DevRep DevRep               -- ^ 'TKProduct'
  | DUnit                             -- ^ @TKScalar Z1@ and empty products

-- | @type role Device nominal@: the sibling of @newtype Concrete y = Concrete (RepConcrete y)@.
newtype Device (y :: TK) = Device { unDevice :: DevRep }

-- | @BoolOf Device@: a host 'Bool' whenever operands are concrete, symbolic inside index functions.
data DevBool = BHost !Bool | BSym BExpr

devScalarHost :: forall r. HbElt r => r -> Device (TKScalar r)
devScalarHost = Device . DHost . toHostScalar

-- | A scalar as a 0-d device array (uploading a host literal when needed: @hb_full@).
devScalarAsArr :: forall r. HbElt r => Device (TKScalar r) -> DevArr
devScalarAsArr (Device rep) = case rep of
  DArr a -> a
  DHost h -> wrapOut $ \out -> with (fromHostScalar @r h) $ \pv ->
               c_hb_full (hbDtype @r) 0 nullPtr (castPtr pv) out
  _ -> error "devScalarAsArr: symbolic scalar escaped an index function"

-- | Forces a scalar to the host: the ONE synchronising call of a training step (the loss), @hb_scalar_to_host@.
devScalarToHost :: forall r. HbElt r => Device (TKScalar r) -> r
devScalarToHost (Device rep) = case rep of
  DHost h -> fromHostScalar h
  DArr a -> unsafePerformIO $ withDev a $ \pa -> alloca $ \pv -> do
    hbCheck (c_hb_scalar_to_host pa (castPtr pv))
    peek pv
  _ -> error "devScalarToHost: symbolic scalar escaped an index function"

deviceInit :: Int -> IO ()
deviceInit ordinal = hbCheck (c_hb_init (fromIntegral ordinal))

deviceSync :: IO ()
deviceSync = hbCheck c_hb_sync
