{-# LANGUAGE AllowAmbiguousTypes, BangPatterns, CPP, DataKinds, FlexibleContexts, FlexibleInstances, GADTs, LambdaCase,
             MultiParamTypeClasses, QuantifiedConstraints, RankNTypes, ScopedTypeVariables, TypeApplications,
             TypeFamilies, TypeOperators, UndecidableInstances #-}
{-# OPTIONS_GHC -Wno-orphans -Wno-missing-methods #-}
-- | @instance BaseTensor Device@: the sibling of "HordeAd.Core.OpsConcrete" (reference OpsConcrete.hs:97-875) whose
-- methods run on a B200 through libhordeb200.  Every method of 'LetTensor', 'ShareTensor', 'BaseTensor'
-- (Ops.hs:157-1255) and 'ConvertTensor' (ConvertTensor.hs:25-228) is defined here.
--
-- * Shaped, ranked and mixed variants share one untyped implementation: the handle carries the run-time shape, so
--   @trsum@, @tssum@ and @txsum@ are the same @hb_sum_leading@ call (ox-arrays' three flavours differ in types only).
-- * Metadata-only operations (transpose, replicate, slice, reverse, most reshapes, index by a concrete prefix) return
--   views; no kernel runs (DESIGN.md section 2).
-- * Index functions of gather / scatter are REIFIED: the closure is applied to symbolic loop variables
--   ('IxVar'), the resulting 'IxExpr's are compiled to the library's register bytecode ('compileIx') and handed to
--   @hb_ixprog_build@; inside the library the program becomes straight-line CUDA C compiled by NVRTC for sm_100a
--   (csrc/jit.cu).  This is what OpsAst.hs:474-485 does for the AST instance, applied one level lower.
-- * Everything above this file -- interpretAstFull, gradientFromDelta (DeltaEval.hs), cgrad, the optimizers -- is
--   polymorphic in @target@ and runs unchanged (SURVEY.md section 8b).
--
-- NOT type-checked here (no GHC in the build image): horde-ad_b200/device.py is the executable image of the same
-- method-to-entry-point mapping and is what tests/ drive; tests/test_haskell_binding.py checks that every class
-- method of the reference is defined below and that every @c_hb_*@ name used exists in the generated FFI module.
module HordeAd.Core.OpsDevice
  ( module HordeAd.Core.CarriersDevice
  , compileIx, reifyIndexFunction, devShape, devToVector, devFromVector
  ) where

import Prelude

import Control.Monad (forM_, zipWithM_)
import Data.Bits (shiftL, (.|.))
import Data.Int (Int64)
import Data.IORef
import Data.List (foldl')
import Data.List.NonEmpty (NonEmpty)
import Data.List.NonEmpty qualified as NonEmpty
import Data.Proxy (Proxy (Proxy))
import Data.Vector qualified as V
import Data.Vector.Storable qualified as VS
import Data.Word (Word8)
import Foreign.C.Types (CDouble (..), CInt (..))
import Foreign.ForeignPtr (ForeignPtr, newForeignPtr_, withForeignPtr)
import Foreign.Marshal.Alloc (alloca, allocaBytes)
import Foreign.Marshal.Array (peekArray, withArray, withArrayLen)
import Foreign.Marshal.Utils (with)
import Foreign.Ptr (Ptr, castPtr, nullPtr, plusPtr)
import Foreign.Storable (peek, poke, pokeByteOff, sizeOf)
import GHC.Exts (IsList (..))
import GHC.TypeLits (KnownNat, natVal)
import System.IO.Unsafe (unsafePerformIO)

import Data.Array.Nested qualified as Nested
import Data.Array.Nested.Mixed.Shape
import Data.Array.Nested.Permutation qualified as Permutation
import Data.Array.Nested.Ranked.Shape
import Data.Array.Nested.Shaped.Shape
import Data.Boolean (Boolean (..), IfB (..))

import HordeAd.Core.CarriersADVal (cfwdOnParams, crevOnParams, crevOnParamsDt)
import HordeAd.Core.CarriersConcrete (Concrete (..), RepConcrete)
import HordeAd.Core.CarriersDevice
import HordeAd.Core.ConvertTensor
import HordeAd.Core.DeviceFFI
import HordeAd.Core.Ops
import HordeAd.Core.TensorKind
import HordeAd.Core.Types

-- ---------------------------------------------------------------------------------------------------------------
-- Untyped layer: one function per entry point family

type instance PrimalOf Device = Device
type instance DualOf Device = DummyDualTarget
type instance PlainOf Device = Device
type instance ShareOf Device = Device
type instance BoolOf Device = DevBool
type instance HFunOf Device x z = Device x -> Device z

arr :: Device y -> DevArr
arr (Device (DArr a)) = a
arr _ = error "Device: tensor expected"

dev :: DevArr -> Device y
dev = Device . DArr

i64s :: [Int] -> [Int64]
i64s = map fromIntegral

devShape :: DevArr -> [Int]
devShape a = unsafePerformIO $ withDev a $ \pa -> alloca $ \pr -> do
  hbCheck (c_hb_rank pa pr)
  r <- fromIntegral <$> peek pr
  allocaBytes (8 * max 1 r) $ \ps -> do
    hbCheck (c_hb_shape pa ps)
    map fromIntegral <$> peekArray r ps

devSize :: DevArr -> Int
devSize = product . devShape

-- | @tsconcrete@ / @trconcrete@ / @txconcrete@: one H2D copy (@hb_from_host@; pinned staging inside the library).
devFromVector :: forall r. HbElt r => [Int] -> VS.Vector r -> DevArr
devFromVector sh v = wrapOut $ \out -> withArrayLen (i64s sh) $ \n ps -> VS.unsafeWith v $ \pv ->
  c_hb_from_host (dtypeCode @r) (fromIntegral n) ps (castPtr pv) out

-- | The only bulk D2H path (@hb_to_host@; synchronises).  Used by @Show@, tests, and checkpointing.
devToVector :: forall r. HbElt r => DevArr -> VS.Vector r
devToVector a = unsafePerformIO $ do
  let n = devSize a
  fp <- VS.unsafeNew n >>= \mv -> VS.unsafeFreeze mv
  VS.unsafeWith fp $ \pv -> withDev a $ \pa -> hbCheck (c_hb_to_host pa (castPtr pv))
  return fp

unary :: CInt -> DevArr -> DevArr
unary op a = wrapOut $ \out -> withDev a $ \pa -> c_hb_unary op pa out

binary :: CInt -> DevArr -> DevArr -> DevArr
binary op a b = wrapOut $ \out -> withDev a $ \pa -> withDev b $ \pb -> c_hb_binary op pa pb out

sumLeading :: Int -> DevArr -> DevArr
sumLeading n a = wrapOut $ \out -> withDev a $ \pa -> c_hb_sum_leading pa (fromIntegral n) out

sumAll, contiguous, reverse1, argMaxLast, argMinLast, relu :: DevArr -> DevArr
sumAll a = wrapOut $ \out -> withDev a $ \pa -> c_hb_sum_all pa out
contiguous a = wrapOut $ \out -> withDev a $ \pa -> c_hb_contiguous pa out
reverse1 a = wrapOut $ \out -> withDev a $ \pa -> c_hb_reverse pa out
argMaxLast a = wrapOut $ \out -> withDev a $ \pa -> c_hb_argmax_last pa out
argMinLast a = wrapOut $ \out -> withDev a $ \pa -> c_hb_argmin_last pa out
relu a = wrapOut $ \out -> withDev a $ \pa -> c_hb_relu pa out

dotAll, dotInner, matmul2, matvecmul, append2, addAcc :: DevArr -> DevArr -> DevArr
dotAll a b = wrapOut $ \out -> withDev a $ \pa -> withDev b $ \pb -> c_hb_dot_all pa pb out
dotInner a b = wrapOut $ \out -> withDev a $ \pa -> withDev b $ \pb -> c_hb_dot_inner pa pb out
matmul2 a b = wrapOut $ \out -> withDev a $ \pa -> withDev b $ \pb -> c_hb_matmul2 pa pb out
matvecmul a b = wrapOut $ \out -> withDev a $ \pa -> withDev b $ \pb -> c_hb_matvecmul pa pb out
append2 a b = wrapOut $ \out -> withDev a $ \pa -> withDev b $ \pb -> c_hb_append pa pb out
-- | @taddTarget@ on tensors: in place when the accumulator is uniquely referenced (DeltaEval.hs:139-148,317-332).
addAcc a b = wrapOut $ \out -> withDev a $ \pa -> withDev b $ \pb -> c_hb_add_inplace_or_new pa pb out

replicateN, reshape :: [Int] -> DevArr -> DevArr
replicateN shm a = wrapOut $ \out -> withDev a $ \pa -> withArrayLen (i64s shm) $ \n ps ->
  c_hb_replicate pa (fromIntegral n) ps out
reshape sh a = wrapOut $ \out -> withDev a $ \pa -> withArrayLen (i64s sh) $ \n ps ->
  c_hb_reshape pa (fromIntegral n) ps out

transpose :: [Int] -> DevArr -> DevArr
transpose perm a = wrapOut $ \out -> withDev a $ \pa -> withArrayLen (map fromIntegral perm) $ \n pp ->
  c_hb_transpose pa (fromIntegral n) pp out

slice :: Int -> Int -> DevArr -> DevArr
slice i n a = wrapOut $ \out -> withDev a $ \pa -> c_hb_slice pa (fromIntegral i) (fromIntegral n) out

-- | Index by a CONCRETE prefix: a view; an out-of-bounds prefix gives zeros (OpsConcrete.hs:1448).
indexPrefix :: [Int] -> DevArr -> DevArr
indexPrefix ix a = wrapOut $ \out -> withDev a $ \pa -> withArrayLen (i64s ix) $ \n pi_ ->
  c_hb_index_prefix pa (fromIntegral n) pi_ out

stack :: [DevArr] -> DevArr
stack [] = error "tsfromVector: empty vector"
stack as = wrapOut $ \out -> withMany as $ \ps -> withArrayLen ps $ \n pp -> c_hb_stack (fromIntegral n) pp out

withMany :: [DevArr] -> ([Ptr HbArray] -> IO a) -> IO a
withMany [] k = k []
withMany (a : as) k = withDev a $ \p -> withMany as (k . (p :))

castTo, floorTo :: CInt -> DevArr -> DevArr
castTo dt a = wrapOut $ \out -> withDev a $ \pa -> c_hb_cast pa dt out
floorTo dt a = wrapOut $ \out -> withDev a $ \pa -> c_hb_floor pa dt out

iota :: CInt -> Int -> DevArr
iota dt n = wrapOut $ \out -> c_hb_iota dt (fromIntegral n) out

oneHot :: [Int] -> [Int] -> DevArr -> DevArr
oneHot shp ix v = wrapOut $ \out -> withDev v $ \pv -> withArrayLen (i64s shp) $ \n ps -> withArray (i64s ix) $ \pix ->
  c_hb_onehot pv (fromIntegral n) ps pix out

compareAll :: CInt -> DevArr -> DevArr -> Bool
compareAll op a b = unsafePerformIO $ withDev a $ \pa -> withDev b $ \pb -> alloca $ \pr -> do
  hbCheck (c_hb_compare_all op pa pb pr)
  (/= 0) <$> peek pr

-- ---------------------------------------------------------------------------------------------------------------
-- Index functions: reification and compilation to the library's bytecode (struct hb_insn, hordeb200.h:265-277)

data Insn = Insn { iOp :: !Word8, iDst :: !Word8, iA :: !Word8, iB :: !Word8, iC :: !Word8, iN :: !Word8
                 , iImm :: Either Int64 Double, iRegs :: [Word8] }

-- op codes of @enum hb_vmop@, in declaration order
vmIVAR, vmICONST, vmIADD, vmISUB, vmIMUL, vmIQUOT, vmIREM, vmINEG, vmIABS, vmISIGNUM, vmIMIN, vmIMAX, vmILEQ, vmILT,
  vmIEQ, vmBAND, vmBOR, vmBNOT, vmSEL, vmFCONST, vmFADD, vmFUN1, vmFLEQ, vmFLT, vmFEQ, vmI2F, vmFLOOR, vmTBL,
  vmARGMAX, vmARGMIN, vmOUT :: Word8
vmIVAR = 0; vmICONST = 1; vmIADD = 2; vmISUB = 3; vmIMUL = 4; vmIQUOT = 5; vmIREM = 6; vmINEG = 7; vmIABS = 8
vmISIGNUM = 9; vmIMIN = 10; vmIMAX = 11; vmILEQ = 12; vmILT = 13; vmIEQ = 14; vmBAND = 15; vmBOR = 16; vmBNOT = 17
vmSEL = 18; vmFCONST = 19; vmFADD = 20; vmFUN1 = 29; vmFLEQ = 30; vmFLT = 31; vmFEQ = 32; vmI2F = 33; vmFLOOR = 34
vmTBL = 35; vmARGMAX = 36; vmARGMIN = 37; vmOUT = 39

-- | Straight-line code, one fresh register per node (programs of the reference's artifacts have a few dozen nodes;
-- the library rejects programs beyond HB_VM_MAX_INSNS / HB_VM_MAX_REGS and the caller falls back to a host loop).
compileIx :: [IxExpr] -> ([Insn], [DevArr])
compileIx outs = unsafePerformIO $ do
  code <- newIORef []
  tens <- newIORef []
  next <- newIORef (0 :: Int)
  let emit i = modifyIORef' code (i :)
      fresh = do { r <- readIORef next; writeIORef next (r + 1); return (fromIntegral r :: Word8) }
      tensor t = do { ts <- readIORef tens; writeIORef tens (ts ++ [t]); return (fromIntegral (length ts) :: Word8) }
      simple op a b c imm = do { d <- fresh; emit (Insn op d a b c 0 imm []); return d }
      goI :: IxExpr -> IO Word8
      goI = \case
        IxVar k -> simple vmIVAR (fromIntegral k) 0 0 (Left 0)
        IxConst n -> simple vmICONST 0 0 0 (Left n)
        IxBinOp o x y -> do
          a <- goI x; b <- goI y
          simple (case o of { IxAdd -> vmIADD; IxSub -> vmISUB; IxMul -> vmIMUL; IxQuot -> vmIQUOT; IxRem -> vmIREM
                            ; IxMin -> vmIMIN; IxMax -> vmIMAX }) a b 0 (Left 0)
        IxUnOp o x -> do
          a <- goI x
          simple (case o of { IxNeg -> vmINEG; IxAbs -> vmIABS; IxSignum -> vmISIGNUM }) a 0 0 (Left 0)
        IxIf c x y -> do { p <- goB c; a <- goI x; b <- goI y; simple vmSEL p a b (Left 0) }
        IxTbl t ix -> table vmTBL t ix
        IxArgMax t ix -> table vmARGMAX t ix
        IxArgMin t ix -> table vmARGMIN t ix
        IxFloor x -> do { a <- goF x; simple vmFLOOR a 0 0 (Left 0) }
      table op t ix = do
        rs <- mapM goI ix
        k <- tensor t
        d <- fresh
        emit (Insn op d k 0 0 (fromIntegral (length rs)) (Left 0) rs)
        return d
      goF :: FxExpr -> IO Word8
      goF = \case
        FxConst x -> simple vmFCONST 0 0 0 (Right x)
        FxTbl t ix -> table vmTBL t ix
        FxFromInt x -> do { a <- goI x; simple vmI2F a 0 0 (Left 0) }
        FxBin o x y -> do { a <- goF x; b <- goF y; simple (vmFADD + fromIntegral o) a b 0 (Left 0) }
        FxUn o x -> do { a <- goF x; simple vmFUN1 a (fromIntegral o) 0 (Left 0) }
      goB :: BExpr -> IO Word8
      goB = \case
        BConst b -> simple vmICONST 0 0 0 (Left (if b then 1 else 0))
        BCmp o x y -> do { a <- goI x; b <- goI y
                         ; simple (case o of { CmpLeq -> vmILEQ; CmpLt -> vmILT; CmpEq -> vmIEQ }) a b 0 (Left 0) }
        BCmpF o x y -> do { a <- goF x; b <- goF y
                          ; simple (case o of { CmpLeq -> vmFLEQ; CmpLt -> vmFLT; CmpEq -> vmFEQ }) a b 0 (Left 0) }
        BAnd x y -> do { a <- goB x; b <- goB y; simple vmBAND a b 0 (Left 0) }
        BOr x y -> do { a <- goB x; b <- goB y; simple vmBOR a b 0 (Left 0) }
        BNot x -> do { a <- goB x; simple vmBNOT a 0 0 (Left 0) }
  forM_ (zip [0 ..] outs) $ \(k, e) -> do
    r <- goI e
    emit (Insn vmOUT 0 r (fromIntegral (k :: Int)) 0 0 (Left 0) [])
  (,) <$> (reverse <$> readIORef code) <*> readIORef tens

pokeInsn :: Ptr HbInsn -> Insn -> IO ()
pokeInsn p (Insn op d a b c n imm regs) = do
  zipWithM_ (pokeByteOff p) [0 ..] [op, d, a, b, c, n, 0, 0]
  case (imm, regs) of
    (_, _ : _) -> zipWithM_ (\k r -> pokeByteOff p (8 + k) r) [0 ..] (take 8 (regs ++ repeat 0))
    (Left i, _) -> pokeByteOff p 8 i
    (Right f, _) -> pokeByteOff p 8 f

-- | Applies the index closure to the symbolic variables @IxVar 0 .. IxVar (|shm| - 1)@ and builds the device program.
reifyIndexFunction :: Int -> Int -> ([Device (TKScalar Int)] -> [Device (TKScalar Int)]) -> (ForeignPtr HbIxProg, [DevArr])
reifyIndexFunction nVars nOut f = unsafePerformIO $ do
  let outs = map toIx (f [Device (DSymI (IxVar k)) | k <- [0 .. nVars - 1]])
      toIx (Device r) = case r of
        DSymI e -> e
        DHost h -> IxConst (fromIntegral (fromHostScalar h :: Int))
        DArr a -> IxTbl a []                         -- a 0-d device integer (e.g. a @kargMax@ result)
        _ -> error "tsgather: index component is not an integer scalar"
      (code, tensors) = compileIx outs
  allocaBytes (16 * max 1 (length code)) $ \pc -> do
    zipWithM_ (\k i -> pokeInsn (pc `plusPtr` (16 * k)) i) [0 ..] code
    withMany tensors $ \pts -> withArrayLen pts $ \nt ppt -> alloca $ \pout -> do
      hbCheck (c_hb_ixprog_build pc (fromIntegral (length code)) (fromIntegral nVars) (fromIntegral nOut) ppt
                                 (fromIntegral nt) 0 pout)
      prog <- peek pout
      fp <- newForeignPtr_ prog        -- freed with the program cache of the library at hb_shutdown
      return (fp, tensors)

gather :: [Int] -> Int -> ([Device (TKScalar Int)] -> [Device (TKScalar Int)]) -> DevArr -> DevArr
gather shm nShp f t =
  let (prog, _keep) = reifyIndexFunction (length shm) nShp f
  in wrapOut $ \out -> withDev t $ \pt -> withForeignPtr prog $ \pp -> withArrayLen (i64s shm) $ \r ps ->
       c_hb_gather pt (fromIntegral r) ps (fromIntegral nShp) pp out

-- | Deterministic: contributions to one destination are added in row-major source order (OpsConcrete.hs:1594-1601).
scatter :: Int -> [Int] -> ([Device (TKScalar Int)] -> [Device (TKScalar Int)]) -> DevArr -> DevArr
scatter nShm shp f t =
  let (prog, _keep) = reifyIndexFunction nShm (length shp) f
  in wrapOut $ \out -> withDev t $ \pt -> withForeignPtr prog $ \pp -> withArrayLen (i64s shp) $ \r ps ->
       c_hb_scatter_add pt (fromIntegral nShm) (fromIntegral r) ps pp out

-- | An index whose components are all host literals or 0-d device integers that we force (the interpreter's
-- concrete @tsindex@); symbolic components only occur inside 'reifyIndexFunction'.
concreteIx :: [Device (TKScalar Int)] -> [Int]
concreteIx = map devScalarToHost

-- ---------------------------------------------------------------------------------------------------------------
-- Scalars: Num and friends.  Host literals fold on the host, symbolic operands build expressions, 0-d device
-- arrays stay on the device (no synchronisation).

liftI2 :: IxBin -> (Int64 -> Int64 -> Int64) -> CInt -> DevRep -> DevRep -> DevRep
liftI2 sym host op x y = case (x, y) of
  (DHost (HInt a), DHost (HInt b)) -> DHost (HInt (host a b))
  (DSymI a, DSymI b) -> DSymI (IxBinOp sym a b)
  (DSymI a, DHost (HInt b)) -> DSymI (IxBinOp sym a (IxConst b))
  (DHost (HInt a), DSymI b) -> DSymI (IxBinOp sym (IxConst a) b)
  (DSymI a, DArr b) -> DSymI (IxBinOp sym a (IxTbl b []))
  (DArr a, DSymI b) -> DSymI (IxBinOp sym (IxTbl a []) b)
  _ -> DArr (binary op (scalarArr @Int x) (scalarArr @Int y))

scalarArr :: forall r. HbElt r => DevRep -> DevArr
scalarArr = devScalarAsArr @r . Device

liftF2 :: forall r. (HbElt r, Fractional r) => (r -> r -> r) -> CInt -> DevRep -> DevRep -> DevRep
liftF2 host op x y = case (x, y) of
  (DHost a, DHost b) -> DHost (toHostScalar (host (fromHostScalar a) (fromHostScalar b)))
  _ -> DArr (binary op (scalarArr @r x) (scalarArr @r y))

instance {-# OVERLAPPING #-} Num (Device (TKScalar Int)) where
  Device a + Device b = Device (liftI2 IxAdd (+) 0 a b)
  Device a - Device b = Device (liftI2 IxSub (-) 1 a b)
  Device a * Device b = Device (liftI2 IxMul (*) 2 a b)
  negate (Device a) = case a of { DHost (HInt n) -> Device (DHost (HInt (negate n))); DSymI e -> Device (DSymI (IxUnOp IxNeg e))
                                ; _ -> dev (unary 0 (scalarArr @Int a)) }
  abs (Device a) = case a of { DHost (HInt n) -> Device (DHost (HInt (abs n))); DSymI e -> Device (DSymI (IxUnOp IxAbs e))
                             ; _ -> dev (unary 1 (scalarArr @Int a)) }
  signum (Device a) = case a of { DHost (HInt n) -> Device (DHost (HInt (signum n))); DSymI e -> Device (DSymI (IxUnOp IxSignum e))
                                ; _ -> dev (unary 2 (scalarArr @Int a)) }
  fromInteger = Device . DHost . HInt . fromInteger

instance IntegralH (Device (TKScalar Int)) where   -- total: a zero divisor returns the dividend (Types.hs quotH/remH)
  quotH (Device a) (Device b) = Device (liftI2 IxQuot (\x y -> if y == 0 then x else quot x y) 7 a b)
  remH (Device a) (Device b) = Device (liftI2 IxRem (\x y -> if y == 0 then x else rem x y) 8 a b)

instance (HbElt r, Fractional r) => Num (Device (TKScalar r)) where
  Device a + Device b = Device (liftF2 @r (+) 0 a b)
  Device a - Device b = Device (liftF2 @r (-) 1 a b)
  Device a * Device b = Device (liftF2 @r (*) 2 a b)
  negate (Device a) = case a of { DHost h -> devScalarHost (negate (fromHostScalar h :: r)); _ -> dev (unary 0 (scalarArr @r a)) }
  abs (Device a) = case a of { DHost h -> devScalarHost (abs (fromHostScalar h :: r)); _ -> dev (unary 1 (scalarArr @r a)) }
  signum (Device a) = case a of { DHost h -> devScalarHost (signum (fromHostScalar h :: r)); _ -> dev (unary 2 (scalarArr @r a)) }
  fromInteger = devScalarHost . (fromInteger :: Integer -> r)

instance (HbElt r, Fractional r) => Fractional (Device (TKScalar r)) where
  Device a / Device b = Device (liftF2 @r (/) 3 a b)
  recip x = 1 / x
  fromRational = devScalarHost . (fromRational :: Rational -> r)

instance (HbElt r, Floating r) => Floating (Device (TKScalar r)) where
  pi = devScalarHost (pi :: r)
  exp = scalarUn @r 4 exp;    log = scalarUn @r 5 log;    sqrt = scalarUn @r 6 sqrt
  sin = scalarUn @r 7 sin;    cos = scalarUn @r 8 cos;    tan = scalarUn @r 9 tan
  asin = scalarUn @r 10 asin; acos = scalarUn @r 11 acos; atan = scalarUn @r 12 atan
  sinh = scalarUn @r 13 sinh; cosh = scalarUn @r 14 cosh; tanh = scalarUn @r 15 tanh
  asinh = scalarUn @r 16 asinh; acosh = scalarUn @r 17 acosh; atanh = scalarUn @r 18 atanh
  Device a ** Device b = Device (liftF2 @r (**) 4 a b)
  logBase (Device a) (Device b) = Device (liftF2 @r logBase 5 a b)

scalarUn :: forall r. HbElt r => CInt -> (r -> r) -> Device (TKScalar r) -> Device (TKScalar r)
scalarUn op host (Device a) = case a of
  DHost h -> devScalarHost (host (fromHostScalar h))
  _ -> dev (unary op (scalarArr @r a))

-- Tensors: the derived Num / Floating instances of Concrete (CarriersConcrete.hs:333-343) as hb_binary / hb_unary
-- (op codes = OpCodeNum1 / OpCode1 / OpCode2 / OpCodeIntegral2 of Ast.hs:707-726, see enum hb_unop / hb_binop).
tensorNum2 :: CInt -> Device y -> Device y -> Device y
tensorNum2 op a b = dev (binary op (arr a) (arr b))
tensorNum1 :: CInt -> Device y -> Device y
tensorNum1 op a = dev (unary op (arr a))

#define TENSOR_INSTANCES(TY, CTX)                                                                          \
instance CTX => Num (Device (TY)) where {                                                                   \
  (+) = tensorNum2 0; (-) = tensorNum2 1; (*) = tensorNum2 2;                                               \
  negate = tensorNum1 0; abs = tensorNum1 1; signum = tensorNum1 2;                                         \
  fromInteger = error "fromInteger: shape unknown (use srepl / rrepl as with Concrete)" };                  \
instance (CTX, Fractional r) => Fractional (Device (TY)) where {                                            \
  (/) = tensorNum2 3; recip = tensorNum1 3;                                                                 \
  fromRational = error "fromRational: shape unknown" };                                                     \
instance (CTX, Floating r) => Floating (Device (TY)) where {                                                \
  pi = error "pi: shape unknown";                                                                           \
  exp = tensorNum1 4; log = tensorNum1 5; sqrt = tensorNum1 6; sin = tensorNum1 7; cos = tensorNum1 8;      \
  tan = tensorNum1 9; asin = tensorNum1 10; acos = tensorNum1 11; atan = tensorNum1 12; sinh = tensorNum1 13; \
  cosh = tensorNum1 14; tanh = tensorNum1 15; asinh = tensorNum1 16; acosh = tensorNum1 17;                 \
  atanh = tensorNum1 18; (**) = tensorNum2 4; logBase = tensorNum2 5 };                                     \
instance (CTX, RealFloat r) => RealFloatH (Device (TY)) where { atan2H = tensorNum2 6 };                    \
instance (CTX, Integral r) => IntegralH (Device (TY)) where { quotH = tensorNum2 7; remH = tensorNum2 8 }

TENSOR_INSTANCES(TKS sh r, HbElt r)
TENSOR_INSTANCES(TKR n r, HbElt r)
TENSOR_INSTANCES(TKX sh r, HbElt r)

-- ---------------------------------------------------------------------------------------------------------------
-- Booleans and comparisons (CarriersConcrete.hs:310-331): ONE host Bool for whole arrays (hb_compare_all); symbolic
-- inside index functions.

instance Boolean DevBool where
  true = BHost True
  false = BHost False
  notB = \case { BHost b -> BHost (not b); BSym e -> BSym (BNot e) }
  BHost a &&* BHost b = BHost (a && b)
  a &&* b = BSym (BAnd (bexpr a) (bexpr b))
  BHost a ||* BHost b = BHost (a || b)
  a ||* b = BSym (BOr (bexpr a) (bexpr b))

bexpr :: DevBool -> BExpr
bexpr = \case { BHost b -> BConst b; BSym e -> e }

cmpScalarI :: IxCmp -> (Int64 -> Int64 -> Bool) -> CInt -> DevRep -> DevRep -> DevBool
cmpScalarI sym host op x y = case (x, y) of
  (DHost (HInt a), DHost (HInt b)) -> BHost (host a b)
  (DSymI _, _) -> BSym (BCmp sym (ixOf x) (ixOf y))
  (_, DSymI _) -> BSym (BCmp sym (ixOf x) (ixOf y))
  _ -> BHost (compareAll op (scalarArr @Int x) (scalarArr @Int y))
 where ixOf = \case { DSymI e -> e; DHost (HInt n) -> IxConst n; DArr a -> IxTbl a []; _ -> error "EqH: integer expected" }

instance {-# OVERLAPPING #-} EqH Device (TKScalar Int) where
  Device a ==. Device b = cmpScalarI CmpEq (==) 0 a b
instance {-# OVERLAPPING #-} OrdH Device (TKScalar Int) where
  Device a <=. Device b = cmpScalarI CmpLeq (<=) 1 a b
instance (HbElt r, Ord r) => EqH Device (TKScalar r) where
  a ==. b = BHost (devScalarToHost a == devScalarToHost b)
instance (HbElt r, Ord r) => OrdH Device (TKScalar r) where
  a <=. b = BHost (devScalarToHost a <= devScalarToHost b)
instance HbElt r => EqH Device (TKS sh r) where { a ==. b = BHost (compareAll 0 (arr a) (arr b)) }
instance HbElt r => OrdH Device (TKS sh r) where { a <=. b = BHost (compareAll 1 (arr a) (arr b)) }
instance HbElt r => EqH Device (TKR n r) where { a ==. b = BHost (compareAll 0 (arr a) (arr b)) }
instance HbElt r => OrdH Device (TKR n r) where { a <=. b = BHost (compareAll 1 (arr a) (arr b)) }
instance HbElt r => EqH Device (TKX sh r) where { a ==. b = BHost (compareAll 0 (arr a) (arr b)) }
instance HbElt r => OrdH Device (TKX sh r) where { a <=. b = BHost (compareAll 1 (arr a) (arr b)) }

-- | @ifH@ / @tcond@: a host branch when the condition is known; a select node when it is symbolic (only scalars can
-- be symbolic, and only inside index functions).
condRep :: DevBool -> DevRep -> DevRep -> DevRep
condRep (BHost b) u v = if b then u else v
condRep (BSym c) u v = DSymI (IxIf c (ix u) (ix v))
 where ix = \case { DSymI e -> e; DHost (HInt n) -> IxConst n; DArr a -> IxTbl a []; _ -> error "ifH: integer branches expected" }

-- ---------------------------------------------------------------------------------------------------------------
-- The classes

instance LetTensor Device where
  ttlet a f = f a                     -- handles are values: sharing is by reference count
  ttletPrimal a f = f a
  ttletPlain a f = f a
  toShare = id
  tunshare = id
  tappend a b = dev (append2 (arr a) (arr b))
  tD _ t _ = t
  -- folds and scans are host loops over views of the leading axis, exactly as for Concrete (OpsConcrete.hs:60-91);
  -- the product fold `sfold (*) 1` and its reverse rule are recognised below the ABI (hb_prod_fold_grad, scan.cu)
  tfold k _nstk _mstk f acc0 es = foldl' f acc0 (unravelDevice k es)
  tscan k nstk _mstk f acc0 es = stackDevice nstk (scanl f acc0 (unravelDevice k es))

unravelDevice :: SNat k -> Device (BuildTensorKind k y) -> [Device y]
unravelDevice k (Device rep) = case rep of
  DPair a b -> zipWith (\(Device x) (Device y) -> Device (DPair x y)) (unravelDevice k (Device a)) (unravelDevice k (Device b))
  DArr a -> [dev (indexPrefix [i] a) | i <- [0 .. fromSNat' k - 1]]
  _ -> error "unravel: tensor expected"

stackDevice :: SingletonTK y -> [Device y] -> Device (BuildTensorKind k y)
stackDevice stk xs = case stk of
  STKProduct s1 s2 -> let (as, bs) = unzip [ (Device a, Device b) | Device (DPair a b) <- xs ]
                      in Device (DPair (unDevice (stackDevice s1 as)) (unDevice (stackDevice s2 bs)))
  STKScalar -> dev (stack [ scalar0d x | x <- xs ])
  _ -> dev (stack (map arr xs))
 where scalar0d (Device r) = case r of { DArr a -> a; _ -> error "tscan: host scalars are stacked by tsfromVectorLinear" }

instance ShareTensor Device where
  tshare = id
  tunpair (Device (DPair a b)) = (Device a, Device b)
  tunpair _ = error "tunpair: product expected"
  tunravelToListShare k _stk = unravelDevice k

instance BaseTensor Device where
  isConcreteInstance = True           -- eager evaluation: the interpreter treats Device exactly like Concrete
  -- ---- shapes: metadata only (hb_rank / hb_shape, `unsafe` calls)
  rshape = listToShR . devShape . arr
  rlength = length . devShape . arr
  rsize = devSize . arr
  rwidth a = case devShape (arr a) of { n : _ -> n; [] -> error "rwidth: rank 0" }
  sshape @sh _ = knownShS @sh
  slength @sh _ = shsLength (knownShS @sh)
  ssize @sh _ = shsSize (knownShS @sh)
  swidth a = case devShape (arr a) of { n : _ -> n; [] -> error "swidth: rank 0" }
  xshape @sh a = shxFromList (knownShX @sh) (devShape (arr a))
  xlength = length . devShape . arr
  xsize = devSize . arr
  xwidth a = case devShape (arr a) of { n : _ -> n; [] -> error "xwidth: rank 0" }
  tsize stk a = case stk of
    STKScalar -> 1
    STKProduct s1 s2 -> let (a1, a2) = tunpair a in tsize s1 a1 + tsize s2 a2
    _ -> devSize (arr a)
  tftk stk a = case stk of
    STKScalar -> FTKScalar
    STKR _ x -> FTKR (rshape a) (ftkElement x)
    STKS sh x -> FTKS sh (ftkElement x)
    STKX _ x -> FTKX (xshape a) (ftkElement x)
    STKProduct s1 s2 -> let (a1, a2) = tunpair a in FTKProduct (tftk s1 a1) (tftk s2 a2)
  tpair (Device a) (Device b) = Device (DPair a b)
  tproject1 = fst . tunpair
  tproject2 = snd . tunpair
  kcond b (Device u) (Device v) = Device (condRep b u v)
  scond b (Device u) (Device v) = Device (condRep b u v)
  tcond _ b (Device u) (Device v) = Device (condRep b u v)
  -- ---- upload (tsconcrete = hb_from_host; scalars stay host literals until a kernel needs them)
  tkconcrete = devScalarHost
  trconcrete a = dev (devFromVector (shrToList (Nested.rshape a)) (Nested.rtoVector a))
  tsconcrete a = dev (devFromVector (shsToList (Nested.sshape a)) (Nested.stoVector a))
  txconcrete a = dev (devFromVector (shxToList (Nested.mshape a)) (Nested.mtoVector a))
  tconcrete ftk (Concrete a) = concreteToDevice ftk a
  -- ---- stacking / unstacking (hb_stack = one gather-free copy kernel; unravel = views)
  trfromVector v = dev (stack (map arr (V.toList v)))
  trfromVectorN shm v = dev (reshapeLeading (shrToList shm) (stack (map arr (V.toList v))))
  trfromVectorLinear shm v = dev (reshape (shrToList shm) (stackScalars (V.toList v)))
  trunravelToList a = [dev (indexPrefix [i] (arr a)) | i <- [0 .. rwidth a - 1]]
  trunravelToListN shm a = [dev (indexPrefix ix (arr a)) | ix <- enumIndices (shrToList shm)]
  trtoListLinear a = [dev (indexPrefix ix (arr a)) | ix <- enumIndices (devShape (arr a))]
  tsfromVector v = dev (stack (map arr (V.toList v)))
  tsfromVectorN shm v = dev (reshapeLeading (shsToList shm) (stack (map arr (V.toList v))))
  tsfromVectorLinear shm v = dev (reshape (shsToList shm) (stackScalars (V.toList v)))
  tsunravelToList a = [dev (indexPrefix [i] (arr a)) | i <- [0 .. swidth a - 1]]
  tsunravelToListN shm a = [dev (indexPrefix ix (arr a)) | ix <- enumIndices (shsToList shm)]
  tstoListLinear a = [dev (indexPrefix ix (arr a)) | ix <- enumIndices (devShape (arr a))]
  tsfromIxR l = dev (stack (map arr (foldr (:) [] l)))
  txfromVector v = dev (stack (map arr (V.toList v)))
  txfromVectorN shm v = dev (reshapeLeading (shxToList shm) (stack (map arr (V.toList v))))
  txfromVectorLinear shm v = dev (reshape (shxToList shm) (stackScalars (V.toList v)))
  txunravelToList a = [dev (indexPrefix [i] (arr a)) | i <- [0 .. xwidth a - 1]]
  txunravelToListN shm a = [dev (indexPrefix ix (arr a)) | ix <- enumIndices (shxToList shm)]
  txtoListLinear a = [dev (indexPrefix ix (arr a)) | ix <- enumIndices (devShape (arr a))]
  -- ---- reductions and contractions (reduce.cu, contract.cu, gemm_dmma.cu, contract_tf32.cu)
  trsum = dev . sumLeading 1 . arr
  trsumN n = dev . sumLeading n . arr
  trsum0 = dev . sumAll . arr
  trdot0 a b = dev (dotAll (arr a) (arr b))
  trdot1In a b = dev (dotInner (arr a) (arr b))
  trmatvecmul m v = dev (matvecmul (arr m) (arr v))
  trmatmul2 a b = dev (matmul2 (arr a) (arr b))
  trreplicate k = dev . replicateN [k] . arr
  trreplicateN shm = dev . replicateN (shrToList shm) . arr
  trreplicate0N @r sh = dev . replicateN (shrToList sh) . devScalarAsArr @r
  tssum = dev . sumLeading 1 . arr
  tssumN @shm = dev . sumLeading (shsLength (knownShS @shm)) . arr
  tssum0 = dev . sumAll . arr
  tsdot0 a b = dev (dotAll (arr a) (arr b))
  tsdot1In _ a b = dev (dotInner (arr a) (arr b))       -- `ssumN (sdot1In a b)` of contracted artifacts: lazy.cu fuses
  tsmatvecmul m v = dev (matvecmul (arr m) (arr v))
  tsmatmul2 a b = dev (matmul2 (arr a) (arr b))
  tsreplicate k = dev . replicateN [fromSNat' k] . arr
  tsreplicateN shm = dev . replicateN (shsToList shm) . arr
  tsreplicate0N @_ @r shm = dev . replicateN (shsToList shm) . devScalarAsArr @r
  txsum = dev . sumLeading 1 . arr
  txsumN n = dev . sumLeading n . arr
  txsum0 = dev . sumAll . arr
  txdot0 a b = dev (dotAll (arr a) (arr b))
  txdot1In _ a b = dev (dotInner (arr a) (arr b))
  txmatvecmul _ _ m v = dev (matvecmul (arr m) (arr v))
  txmatmul2 a b = dev (matmul2 (arr a) (arr b))
  txreplicate k = dev . replicateN [fromSNat' k] . arr
  txreplicateN shm = dev . replicateN (shxToList shm) . arr
  txreplicate0N @r sh = dev . replicateN (shxToList sh) . devScalarAsArr @r
  -- ---- indexing, gather, scatter (ixprog.cu, jit.cu); out-of-bounds: gather/index -> zeros, scatter/oneHot -> dropped
  trindex a ix = indexOrGather a (foldr (:) [] ix)
  trindex0 a ix = indexOrGather a (foldr (:) [] ix)
  troneHot sh v ix = dev (oneHot (shrToList sh) (concreteIx (foldr (:) [] ix)) (arr v))
  trscatter @m @_ @p sh v f = dev (scatter (valueOf @m) (take (valueOf @p) (shrToList sh)) (ixFun f) (arr v))
  trscatter1 @_ @p sh v f = dev (scatter 1 (take (valueOf @p) (shrToList sh)) (ixFun (f . headIx)) (arr v))
  trgather @m @_ @p sh v f = dev (gather (take (valueOf @m) (shrToList sh)) (valueOf @p) (ixFun f) (arr v))
  trgather1 @_ @p k v f = dev (gather [k] (valueOf @p) (ixFun (f . headIx)) (arr v))
  tsindex a ix = indexOrGather a (foldr (:) [] ix)
  tsindex0 a ix = indexOrGather a (foldr (:) [] ix)
  tsoneHot @sh1 v ix = dev (oneHot (shsToList (knownShS @sh1)) (concreteIx (foldr (:) [] ix)) (arr v))
  tsscatter @shm @_ @shp v f =
    dev (scatter (shsLength (knownShS @shm)) (shsToList (knownShS @shp)) (ixFun f) (arr v))
  tsscatter1 @_ @_ @shp v f = dev (scatter 1 (shsToList (knownShS @shp)) (ixFun (f . headIx)) (arr v))
  tsgather @shm @_ @shp v f =
    dev (gather (shsToList (knownShS @shm)) (shsLength (knownShS @shp)) (ixFun f) (arr v))
  tsgather1 @n2 @_ @shp v f =
    dev (gather [fromIntegral (natVal (Proxy @n2))] (shsLength (knownShS @shp)) (ixFun (f . headIx)) (arr v))
  txindex a ix = indexOrGather a (foldr (:) [] ix)
  txindex0 a ix = indexOrGather a (foldr (:) [] ix)
  txoneHot sh v ix = dev (oneHot (shxToList sh) (concreteIx (foldr (:) [] ix)) (arr v))
  txscatter @shm @_ @shp sh v f = dev (scatter (ssxLength (knownShX @shm)) (take (ssxLength (knownShX @shp)) (shxToList sh)) (ixFun f) (arr v))
  txscatter1 @_ @_ @shp sh v f = dev (scatter 1 (take (ssxLength (knownShX @shp)) (shxToList sh)) (ixFun (f . headIx)) (arr v))
  txgather @shm @_ @shp sh v f = dev (gather (take (ssxLength (knownShX @shm)) (shxToList sh)) (ssxLength (knownShX @shp)) (ixFun f) (arr v))
  txgather1 @_ @_ @shp k v f = dev (gather [fromSNat' k] (ssxLength (knownShX @shp)) (ixFun (f . headIx)) (arr v))
  -- ---- casts, floor, iota, argmin / argmax (ties: first extremum in row-major order)
  tkfloor @r @r2 x = case unDevice x of
    DSymF e -> Device (DSymI (IxFloor e))
    _ -> dev (floorTo (dtypeCode @r2) (devScalarAsArr @r x))
  tkfromIntegral @r1 @r2 x = case unDevice x of
    DSymI e -> Device (DSymF (FxFromInt e))
    DHost h -> devScalarHost (fromIntegral (fromHostScalar h :: r1) :: r2)
    _ -> dev (castTo (dtypeCode @r2) (devScalarAsArr @r1 x))
  tkcast @r1 @r2 x = case unDevice x of
    DHost h -> devScalarHost (realToFrac (fromHostScalar h :: r1) :: r2)
    _ -> dev (castTo (dtypeCode @r2) (devScalarAsArr @r1 x))
  -- a 0-d device integer, usable as an index component without a host round trip; on a symbolically indexed
  -- window (the max-pool of the vectorised artifacts: `kargMax (t !$ [i, j, ...])`) the symbolic form
  tkargMin a = case unDevice a of { DSymT t ix -> Device (DSymI (IxArgMin t ix)); _ -> dev (argMinLast (arr a)) }
  tkargMax a = case unDevice a of { DSymT t ix -> Device (DSymI (IxArgMax t ix)); _ -> dev (argMaxLast (arr a)) }
  trfloor @_ @r2 = dev . floorTo (dtypeCode @r2) . arr
  trfromIntegral @_ @r2 = dev . castTo (dtypeCode @r2) . arr
  trcast @_ @r2 = dev . castTo (dtypeCode @r2) . arr
  trargMin = dev . argMinLast . arr
  trargMax = dev . argMaxLast . arr
  triota @r n = dev (iota (dtypeCode @r) n)
  tsfloor @_ @r2 = dev . floorTo (dtypeCode @r2) . arr
  tsfromIntegral @_ @r2 = dev . castTo (dtypeCode @r2) . arr
  tscast @_ @r2 = dev . castTo (dtypeCode @r2) . arr
  tsargMin = dev . argMinLast . arr
  tsargMax = dev . argMaxLast . arr
  tsiota @n @r = dev (iota (dtypeCode @r) (fromIntegral (natVal (Proxy @n))))
  txfloor @_ @r2 = dev . floorTo (dtypeCode @r2) . arr
  txfromIntegral @_ @r2 = dev . castTo (dtypeCode @r2) . arr
  txcast @_ @r2 = dev . castTo (dtypeCode @r2) . arr
  txargMin = dev . argMinLast . arr
  txargMax = dev . argMaxLast . arr
  txiota @n @r = dev (iota (dtypeCode @r) (fromIntegral (natVal (Proxy @n))))
  -- ---- structural operations: views, except append and non-collapsible reshapes (array.cu)
  trappend a b = dev (append2 (arr a) (arr b))
  trconcat = dev . foldr1 append2 . map arr . NonEmpty.toList
  trslice i n = dev . slice i n . arr
  trreverse = dev . reverse1 . arr
  trtranspose perm = dev . transpose perm . arr
  trreshape sh = dev . reshape (shrToList sh) . arr
  tsappend a b = dev (append2 (arr a) (arr b))
  tsslice i n _ = dev . slice (fromSNat' i) (fromSNat' n) . arr
  tsreverse = dev . reverse1 . arr
  tstranspose perm = dev . transpose (Permutation.permToList' perm) . arr
  tsreshape sh2 = dev . reshape (shsToList sh2) . arr
  txappend a b = dev (append2 (arr a) (arr b))
  txconcat = dev . foldr1 append2 . map arr . NonEmpty.toList
  txslice i n _ = dev . slice (fromSNat' i) (fromSNat' n) . arr
  txreverse = dev . reverse1 . arr
  txtranspose perm = dev . transpose (Permutation.permToList' perm) . arr
  txreshape sh2 = dev . reshape (shxToList sh2) . arr
  -- ---- builds and maps: host loops (only reached by code the vectoriser did not touch -- "must work, not be fast",
  --      SURVEY.md 8a-11); each element is a C-ABI call, results are stacked by one kernel
  tkbuild1 @k f = dev (stackScalars [f (Device (hostInt i)) | i <- [0 .. fromIntegral (natVal (Proxy @k)) - 1]])
  tkbuild @sh f = dev (reshape (shsToList (knownShS @sh))
                        (stackScalars [f (ixsFromHost ix) | ix <- enumIndices (shsToList (knownShS @sh))]))
  trbuild1 k f = dev (stack [arr (f (Device (hostInt i))) | i <- [0 .. k - 1]])
  trbuild sh f = dev (reshapeLeading (shrToList sh) (stack [arr (f (ixrFromHost ix)) | ix <- enumIndices (shrToList sh)]))
  trmap0N f a = dev (reshape (devShape (arr a)) (stackScalars [f (dev (indexPrefix ix (arr a))) | ix <- enumIndices (devShape (arr a))]))
  trzipWith0N f a b = dev (reshape (devShape (arr a)) (stackScalars
    [f (dev (indexPrefix ix (arr a))) (dev (indexPrefix ix (arr b))) | ix <- enumIndices (devShape (arr a))]))
  tsbuild1 @k f = dev (stack [arr (f (Device (hostInt i))) | i <- [0 .. fromIntegral (natVal (Proxy @k)) - 1]])
  tsbuild @shm f = dev (reshapeLeading (shsToList (knownShS @shm))
                         (stack [arr (f (ixsFromHost ix)) | ix <- enumIndices (shsToList (knownShS @shm))]))
  tsmap0N f a = dev (reshape (devShape (arr a)) (stackScalars [f (dev (indexPrefix ix (arr a))) | ix <- enumIndices (devShape (arr a))]))
  tszipWith0N f a b = dev (reshape (devShape (arr a)) (stackScalars
    [f (dev (indexPrefix ix (arr a))) (dev (indexPrefix ix (arr b))) | ix <- enumIndices (devShape (arr a))]))
  txbuild1 @k f = dev (stack [arr (f (Device (hostInt i))) | i <- [0 .. fromIntegral (natVal (Proxy @k)) - 1]])
  txbuild sh f = dev (reshapeLeading (shxToList sh) (stack [arr (f (ixxFromHost ix)) | ix <- enumIndices (shxToList sh)]))
  tbuild1 k stk f = stackDevice stk [f (Device (hostInt i)) | i <- [0 .. fromSNat' k - 1]]
  -- ---- mapAccum: a host loop over leading-axis views, strict in the accumulator (OpsConcrete.hs:990-1019); the
  --      derivative arguments are unused by an eager target, as in the Concrete instance
  tmapAccumRDer proxy k accftk bftk eftk f df rf acc0 es =
    let (acc, bs) = tunpair (tmapAccumLDer proxy k accftk bftk eftk f df rf acc0 (treverse k (ftkToSTK eftk) es))
    in tpair acc (treverse k (ftkToSTK bftk) bs)
  tmapAccumLDer _ k _accftk bftk _eftk f _df _rf acc0 es =
    let step (!acc, bs) e = let (acc', b) = tunpair (f (tpair acc e)) in (acc', b : bs)
        (accN, bsRev) = foldl' step (acc0, []) (unravelDevice k es)
    in tpair accN (stackDevice (ftkToSTK bftk) (reverse bsRev))
  tmapAccumR proxy k accftk bftk eftk f = tmapAccumRDer proxy k accftk bftk eftk (tlambda (FTKProduct accftk eftk) (HFun (\x -> let (a, e) = tunpair x in f a e))) undefined undefined
  tmapAccumL proxy k accftk bftk eftk f = tmapAccumLDer proxy k accftk bftk eftk (tlambda (FTKProduct accftk eftk) (HFun (\x -> let (a, e) = tunpair x in f a e))) undefined undefined
  tapply f = f
  tlambda _ f = unHFun f
  -- derivatives of closures: the dual-number machinery over THIS carrier (ADVal Device), exactly as for Concrete
  tgrad xftk h = snd . crevOnParams Nothing (unHFun h) xftk
  tvjp xftk h db_a = let (db, a) = tunpair db_a in snd (crevOnParamsDt db (unHFun h) xftk a)
  tjvp xftk h da_a = let (da, a) = tunpair da_a in snd (cfwdOnParams xftk a (unHFun h) da)
  tprimalPart = id
  tdualPart stk t = DummyDualTarget (tftk stk t)
  tplainPart = id
  tfromPrimal _ t = t
  tfromDual (DummyDualTarget ftk) = tdefTarget ftk
  tfromPlain _ t = t
  tScale _ _ t = t
  tsum k stk u = case stk of
    STKProduct s1 s2 -> let (u1, u2) = tunpair u in tpair (tsum k s1 u1) (tsum k s2 u2)
    _ -> dev (sumLeading 1 (arr u))
  treplicate k stk u = case stk of
    STKProduct s1 s2 -> let (u1, u2) = tunpair u in tpair (treplicate k s1 u1) (treplicate k s2 u2)
    STKScalar -> dev (replicateN [fromSNat' k] (scalarArrAny u))
    _ -> dev (replicateN [fromSNat' k] (arr u))
  treverse k stk u = case stk of
    STKProduct s1 s2 -> let (u1, u2) = tunpair u in tpair (treverse k s1 u1) (treverse k s2 u2)
    _ -> dev (reverse1 (arr u))
  -- ---- cotangent accumulation (DeltaEval.hs:139-148,317-332): in place when uniquely referenced
  taddTarget stk a b = case stk of
    STKProduct s1 s2 -> let (a1, a2) = tunpair a; (b1, b2) = tunpair b in tpair (taddTarget s1 a1 b1) (taddTarget s2 a2 b2)
    STKScalar -> scalarAdd a b
    _ -> dev (addAcc (arr a) (arr b))
  tmultTarget stk a b = case stk of
    STKProduct s1 s2 -> let (a1, a2) = tunpair a; (b1, b2) = tunpair b in tpair (tmultTarget s1 a1 b1) (tmultTarget s2 a2 b2)
    STKScalar -> scalarMul a b
    _ -> dev (binary 2 (arr a) (arr b))
  tsum0Target ftk a = case ftk of
    FTKProduct f1 f2 -> let (a1, a2) = tunpair a in tsum0Target f1 a1 + tsum0Target f2 a2
    FTKScalar -> dev (castTo 0 (scalarArrAny a))
    _ -> dev (castTo 0 (sumAll (arr a)))
  tdot0Target ftk a b = case ftk of
    FTKProduct f1 f2 -> let (a1, a2) = tunpair a; (b1, b2) = tunpair b in tdot0Target f1 a1 b1 + tdot0Target f2 a2 b2
    FTKScalar -> dev (castTo 0 (binary 2 (scalarArrAny a) (scalarArrAny b)))
    _ -> dev (castTo 0 (dotAll (arr a) (arr b)))
  xmcast _ = dev . arr                -- a change of the static shape annotation only

instance ConvertTensor Device where
  -- conversions between scalar / ranked / shaped / mixed are re-taggings of one handle (ConvertTensor.hs:36-69); a
  -- scalar converted to a tensor is uploaded as a 0-d array, a 0-d tensor converted to a scalar stays on the device
  tconvert _ _ (Device r) = Device r
  kfromR = dev . arr;  kfromS = dev . arr;  kfromX = dev . arr
  rfromK = dev . scalarArrAny;  sfromK = dev . scalarArrAny;  xfromK = dev . scalarArrAny
  rfromS = dev . arr;  rfromX = dev . arr
  sfromR = dev . arr;  sfromX = dev . arr
  xfromR = dev . arr;  xfromS = dev . arr
  -- products of arrays <-> arrays of products: struct-of-arrays on the device, so zip/unzip are re-pairings
  rzip p = let (a, b) = tunpair p in Device (DPair (unDevice a) (unDevice b))
  runzip (Device (DPair a b)) = tpair (Device a) (Device b)
  runzip _ = error "runzip: product expected"
  szip p = let (a, b) = tunpair p in Device (DPair (unDevice a) (unDevice b))
  sunzip (Device (DPair a b)) = tpair (Device a) (Device b)
  sunzip _ = error "sunzip: product expected"
  xzip p = let (a, b) = tunpair p in Device (DPair (unDevice a) (unDevice b))
  xunzip (Device (DPair a b)) = tpair (Device a) (Device b)
  xunzip _ = error "xunzip: product expected"
  -- nesting is a type-level regrouping of the axes of one dense array: no device work (DESIGN.md section 0)
  rnest _ = dev . arr;   rnestS _ = dev . arr;  rnestX _ = dev . arr
  snestR _ = dev . arr;  snest _ = dev . arr;   snestX _ = dev . arr
  xnestR _ = dev . arr;  xnestS _ = dev . arr;  xnest _ = dev . arr
  runNest = dev . arr;   runNestS = dev . arr;  runNestX = dev . arr
  sunNestR = dev . arr;  sunNest = dev . arr;   sunNestX = dev . arr
  xunNestR = dev . arr;  xunNestS = dev . arr;  xunNest = dev . arr
  tpairConv = tpair
  tunpairConv = tunpair

-- ---------------------------------------------------------------------------------------------------------------
-- Helpers of the instances

scalarArrAny :: Device y -> DevArr
scalarArrAny (Device r) = case r of
  DArr a -> a
  DHost (HDouble x) -> devScalarAsArr (devScalarHost x)
  DHost (HFloat x) -> devScalarAsArr (devScalarHost x)
  DHost (HInt n) -> devScalarAsArr (devScalarHost n)
  DHost (HBool b) -> devScalarAsArr (devScalarHost b)
  _ -> error "scalar expected"

scalarAdd, scalarMul :: Device y -> Device y -> Device y
scalarAdd a b = dev (binary 0 (scalarArrAny a) (scalarArrAny b))
scalarMul a b = dev (binary 2 (scalarArrAny a) (scalarArrAny b))

stackScalars :: [Device (TKScalar r)] -> DevArr
stackScalars = stack . map scalarArrAny

reshapeLeading :: [Int] -> DevArr -> DevArr
reshapeLeading shm a = case devShape a of
  _ : rest -> reshape (shm ++ rest) a
  [] -> a

enumIndices :: [Int] -> [[Int]]
enumIndices = mapM (\n -> [0 .. n - 1])

ixsFromHost :: [Int] -> IxSOf Device sh
ixsFromHost = fromList . map (Device . hostInt)
ixrFromHost :: [Int] -> IxROf Device n
ixrFromHost = fromList . map (Device . hostInt)
ixxFromHost :: [Int] -> IxXOf Device sh
ixxFromHost = fromList . map (Device . hostInt)

-- | @tsgather1 v f = tsgather v (\(i :.$ _) -> f i)@: the one-variable closure as a closure on the index list
headIx :: (IsList ix, Item ix ~ Device (TKScalar Int)) => ix -> Device (TKScalar Int)
headIx = head . toList

-- | The index closure on plain lists of symbolic scalars (IxR / IxS / IxX are 'IsList's of their components).
ixFun :: (IsList ixm, Item ixm ~ Device (TKScalar Int), IsList ixp, Item ixp ~ Device (TKScalar Int))
      => (ixm -> ixp) -> [Device (TKScalar Int)] -> [Device (TKScalar Int)]
ixFun f = toList . f . fromList

valueOf :: forall n. KnownNat n => Int
valueOf = fromIntegral (natVal (Proxy @n))

-- | A concrete index is a view; an index with 0-d device components is a rank-0 gather (no host round trip); an
-- index with SYMBOLIC components (we are inside the reification of an enclosing index function) yields a symbolic
-- value: a table look-up for a full index, a symbolically indexed window otherwise.
indexOrGather :: Device y -> [Device (TKScalar Int)] -> Device z
indexOrGather a ix
  | all isHost ix = dev (indexPrefix (concreteIx ix) (arr a))
  | any isSym ix =
      let es = map ixOf ix
          t = arr a
      in if length ix == length (devShape t)
           then Device (if devIsIntegral t then DSymI (IxTbl t es) else DSymF (FxTbl t es))
           else Device (DSymT t es)
  | otherwise = dev (gather [] (length ix) (const ix) (arr a))
 where
  isHost (Device r) = case r of { DHost _ -> True; _ -> False }
  isSym (Device r) = case r of { DSymI _ -> True; _ -> False }
  ixOf (Device r) = case r of
    DSymI e -> e
    DHost (HInt n) -> IxConst n
    DArr t -> IxTbl t []
    _ -> error "tsindex: integer index expected"

devIsIntegral :: DevArr -> Bool
devIsIntegral a = unsafePerformIO $ withDev a $ \pa -> alloca $ \pd -> do
  hbCheck (c_hb_dtype_of pa pd)
  (>= 2) <$> peek pd                    -- hb_dtype: 0 = f64, 1 = f32, then the integral types and bool

concreteToDevice :: FullShapeTK y -> RepConcrete y -> Device y
concreteToDevice ftk a = case ftk of
  FTKScalar -> devScalarHost a
  FTKR{} -> trconcrete a
  FTKS{} -> tsconcrete a
  FTKX{} -> txconcrete a
  FTKProduct f1 f2 -> tpair (concreteToDevice f1 (fst a)) (concreteToDevice f2 (snd a))
