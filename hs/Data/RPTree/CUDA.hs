{-# LANGUAGE ForeignFunctionInterface #-}
{-# LANGUAGE BangPatterns #-}
-- | Drop-in GPU back end for the hot path of "Data.RPTree" (rp-tree-0.7.1):
--   'forestBatch' / 'forest' (streaming, any chunk size) -> 'candidates' / 'knn' / 'knnPQ' / 'knnH' -> 'recallWith',
--   specialised to 'Double' data ('DVector' here; 'SVector' points go through 'setPointsSparse') and 'metricL2'.
--
-- Every function keeps the signature shape of its namesake in src/Data/RPTree.hs / Batch.hs / Conduit.hs and marshals to
-- librpforest.so (include/rpforest.h).  The hyperplanes are drawn HERE, with the library's own
-- @sample seed (replicateM ntrees (V.replicateM maxd (sparse pnz dim stdNormal)))@ (Batch.hs:59-61), and handed to
-- the engine through @rpf_set_hyperplanes@, so trees are bit-identical to the pure implementation by construction.
--
-- NOTE: written against the C ABI but NOT compiled in the authoring environment (no GHC there).
module Data.RPTree.CUDA
  ( GpuForest, forestBatch, treeBatch, forest, forestStreaming, forestOn, knn, knnPQ, knnH, candidates, recallWith, toRPForest
  , withDevice, withDevices, saveForest, setPointsSparse
  ) where

import Control.Exception (throwIO, ErrorCall(..))
import Control.Monad (replicateM, when, forM, forM_)
import Control.Monad.IO.Class (MonadIO(..))
import Data.Int (Int32, Int64)
import Data.Word (Word32, Word64)
import Foreign.C.String (CString, peekCString, withCString)
import Foreign.C.Types (CInt(..), CDouble(..))
import Foreign.ForeignPtr (ForeignPtr, newForeignPtr, withForeignPtr)
import Foreign.Marshal.Alloc (alloca)
import Foreign.ForeignPtr (mallocForeignPtrArray)
import Foreign.Marshal.Array (allocaArray, peekArray, withArrayLen, withArray)
import Foreign.Ptr (Ptr, FunPtr, nullPtr)
import Foreign.Storable (peek)
import System.IO.Unsafe (unsafePerformIO)

import qualified Data.Conduit as C
import qualified Data.Conduit.Combinators as CC
import qualified Data.Conduit.List as CL (chunksOf)
import qualified Data.IntMap.Strict as IM
import qualified Data.Vector as V
import qualified Data.Vector.Storable as VS
import qualified Data.Vector.Storable.Mutable as VSM
import qualified Data.Vector.Unboxed as VU
import System.Random.SplitMix.Distributions (sample, stdNormal)

import Data.RPTree.Gen (sparse)
import Data.RPTree.Internal (RPTree(..), RPT(..), RPForest, Embed(..), DVector(..), SVector(..), Margin(..), metricL2)
import Data.Semigroup (Max(..), Min(..))

data RpfHandle

foreign import ccall safe "rpf_create"            c_create   :: Ptr (Ptr RpfHandle) -> CInt -> IO CInt
-- | ONE handle over several GPUs of this process: trees sharded in contiguous blocks, data replicated, NCCL owned by the
-- engine; every entry point below keeps its meaning (rpforest.h, "multi-GPU").
foreign import ccall safe "rpf_create_multi"      c_createM  :: Ptr (Ptr RpfHandle) -> Ptr CInt -> CInt -> IO CInt
foreign import ccall safe "&rpf_destroy"          p_destroy  :: FunPtr (Ptr RpfHandle -> IO ())
foreign import ccall safe "rpf_last_error"        c_lastErr  :: Ptr RpfHandle -> IO CString
foreign import ccall safe "rpf_set_points"        c_setPts   :: Ptr RpfHandle -> Ptr CDouble -> Int64 -> Int32 -> IO CInt
foreign import ccall safe "rpf_set_hyperplanes"   c_setHp    :: Ptr RpfHandle -> Int32 -> Int32 -> Ptr Int64 -> Ptr Int32 -> Ptr CDouble -> IO CInt
foreign import ccall safe "rpf_build"             c_build    :: Ptr RpfHandle -> Int32 -> Int32 -> IO CInt
foreign import ccall safe "rpf_build_from_host"   c_buildH   :: Ptr RpfHandle -> Ptr CDouble -> Int64 -> Int32 -> Int32 -> Int32 -> IO CInt
foreign import ccall safe "rpf_build_chunked"     c_buildCh  :: Ptr RpfHandle -> Int32 -> Int32 -> Int64 -> IO CInt
-- | The fold of Conduit.hs:157-176 driven from Haskell: one 'c_insChunk' == one insertMulti (Internal.hs:243-255); the forest is
-- complete and queryable after every chunk, n is not known in advance.
foreign import ccall safe "rpf_insert_begin"      c_insBegin :: Ptr RpfHandle -> Int32 -> Int32 -> Int32 -> IO CInt
foreign import ccall safe "rpf_insert_chunk"      c_insChunk :: Ptr RpfHandle -> Ptr CDouble -> Int64 -> IO CInt
foreign import ccall safe "rpf_insert_end"        c_insEnd   :: Ptr RpfHandle -> IO CInt
foreign import ccall safe "rpf_num_nodes"         c_numNodes :: Ptr RpfHandle -> IO Int64
foreign import ccall safe "rpf_topology"          c_topology :: Ptr RpfHandle -> Ptr Int64 -> Ptr Int32 -> Ptr Int64 -> Ptr Int64 -> IO CInt
foreign import ccall safe "rpf_tree_export"       c_export   :: Ptr RpfHandle -> Int32 -> Ptr CDouble -> Ptr CDouble -> Ptr CDouble -> Ptr Word32 -> IO CInt
foreign import ccall safe "rpf_forest_export"     c_exportAll :: Ptr RpfHandle -> Ptr CDouble -> Ptr CDouble -> Ptr CDouble -> Ptr Word32 -> IO CInt
-- | Export sink: pinned result buffers registered BEFORE 'c_buildH'; the engine streams thr/mlo/mhi/perm into them while
-- the bottom phase still runs, and 'c_exportAll' with the same pointers only waits (25.1 instead of 27.1 ms end to end
-- at 1M x 128, 32 trees).  'toRPForest' reads the whole forest, so a host that always converts should register a sink.
foreign import ccall safe "rpf_set_export_sink"   c_setSink  :: Ptr RpfHandle -> Ptr CDouble -> Ptr CDouble -> Ptr CDouble -> Ptr Word32 -> IO CInt
foreign import ccall safe "rpf_candidates_count"  c_candCnt  :: Ptr RpfHandle -> Ptr CDouble -> Int64 -> Int32 -> Ptr Int64 -> IO CInt
foreign import ccall safe "rpf_candidates"        c_cand     :: Ptr RpfHandle -> Ptr CDouble -> Int64 -> Int32 -> Ptr Int64 -> Ptr Word32 -> IO CInt
foreign import ccall safe "rpf_knn"               c_knn      :: Ptr RpfHandle -> Ptr CDouble -> Int64 -> Int32 -> Int32 -> Ptr CDouble -> Ptr Word32 -> Ptr Int32 -> IO CInt
foreign import ccall safe "rpf_recall"            c_recall   :: Ptr RpfHandle -> Ptr CDouble -> Int64 -> Int32 -> Ptr CDouble -> IO CInt
foreign import ccall safe "rpf_knn_h_capacity"    c_knnHCap  :: Ptr RpfHandle -> Int32 -> IO Int64
foreign import ccall safe "rpf_knn_h"             c_knnH     :: Ptr RpfHandle -> Ptr CDouble -> Ptr Int32 -> Int64 -> Int32 -> Int64 -> Ptr CDouble -> Ptr Word32 -> Ptr Int32 -> IO CInt
foreign import ccall safe "rpf_set_points_sparse" c_setPtsS  :: Ptr RpfHandle -> Int64 -> Int32 -> Ptr Int64 -> Ptr Int32 -> Ptr CDouble -> IO CInt
foreign import ccall safe "rpf_forest_save"       c_save     :: Ptr RpfHandle -> CString -> Int32 -> IO CInt

-- | A forest living on one B200 -- or sharded over several ('withDevices' / 'forestOn').  The payloads stay on the Haskell side, addressed by row number.
data GpuForest x = GpuForest
  { gfHandle  :: !(ForeignPtr RpfHandle)
  , gfRows    :: !(V.Vector (Embed DVector Double x))   -- ^ row id -> original item
  , gfVectors :: !(IM.IntMap (V.Vector (SVector Double))) -- ^ rvss, as drawn by 'sample'
  , gfNTrees  :: !Int
  , gfMaxD    :: !Int
  }

withDevice :: Int -> (ForeignPtr RpfHandle -> IO a) -> IO a
withDevice dev = withDevices [dev]

-- | An engine over the listed GPUs (one entry: 'rpf_create'; several: 'rpf_create_multi').  'createMulti' is a map over
-- the IntMap of trees (Internal.hs:234-240) and 'knn' folds the trees' candidates (RPTree.hs:176), so the engine gives
-- GPU r a contiguous block of trees and returns exactly the single-GPU results.
withDevices :: [Int] -> (ForeignPtr RpfHandle -> IO a) -> IO a
withDevices devs k = alloca $ \pp -> withArrayLen (map fromIntegral devs) $ \nd pd -> do
  rc <- if nd == 1 then c_create pp (fromIntegral (head devs)) else c_createM pp pd (fromIntegral nd)
  when (rc /= 0) $ throwIO (ErrorCall "rpf_create: no usable CUDA device (there is no CPU fallback)")
  h <- peek pp >>= newForeignPtr p_destroy
  k h

check :: Ptr RpfHandle -> String -> CInt -> IO ()
check h what rc = when (rc /= 0) $ do
  msg <- c_lastErr h >>= peekCString
  throwIO (ErrorCall (what ++ ": " ++ msg))

-- | One row-major n x d 'Storable' buffer out of the unpinned VU.Vectors (Internal.hs:122): written in place, one
-- 'VU.copy'-style pass per row (no intermediate lists; CDouble and Double share their representation).
packRows :: Int -> V.Vector (Embed DVector Double x) -> VS.Vector CDouble
packRows dim xs = VS.create $ do
  buf <- VSM.unsafeNew (V.length xs * dim)
  V.iforM_ xs $ \i xe -> do
    let row = dvVec (eEmbed xe)
    forM_ [0 .. min dim (VU.length row) - 1] $ \j -> VSM.unsafeWrite buf (i * dim + j) (realToFrac (VU.unsafeIndex row j))
  pure buf

-- | CSR over (tree-major, level-minor) of rvss; SVector's VU.Vector (Int, Double) is already SoA (Internal.hs:92-93).
csrOf :: IM.IntMap (V.Vector (SVector Double)) -> ([Int64], [Int32], [CDouble])
csrOf rvss = (scanl (+) 0 (map (fromIntegral . VU.length . svVec) svs), concatMap (map (fromIntegral . fst) . VU.toList . svVec) svs
             , concatMap (map (realToFrac . snd) . VU.toList . svVec) svs)
  where svs = concatMap V.toList (IM.elems rvss)

buildWith :: [Int] -> Maybe Int -> Word64 -> Int -> Int -> Int -> Double -> Int -> V.Vector (Embed DVector Double x) -> IO (GpuForest x)
buildWith devs chunk seed maxd minl ntrees pnz dim xs = withDevices devs $ \fh -> withForeignPtr fh $ \h -> do
  let rvss = sample seed $ do                                   -- Batch.hs:59-61 / Conduit.hs:116-118, verbatim
        rvs <- replicateM ntrees $ V.replicateM maxd (sparse pnz dim stdNormal)
        pure $ IM.fromList $ zip [0 ..] rvs
      (off, idx, val) = csrOf rvss
      buf = packRows dim xs
  withArray off $ \po -> withArray idx $ \pi' -> withArray val $ \pv ->
    c_setHp h (fromIntegral ntrees) (fromIntegral maxd) po pi' pv >>= check h "rpf_set_hyperplanes"
  case chunk of
    -- batch: one call uploads the rows in blocks and projects them as they arrive
    Nothing -> VS.unsafeWith buf $ \p ->
                 c_buildH h p (fromIntegral (V.length xs)) (fromIntegral dim) (fromIntegral maxd) (fromIntegral minl) >>= check h "rpf_build_from_host"
    Just c  -> do
      VS.unsafeWith buf $ \p -> c_setPts h p (fromIntegral (V.length xs)) (fromIntegral dim) >>= check h "rpf_set_points"
      c_buildCh h (fromIntegral maxd) (fromIntegral minl) (fromIntegral c) >>= check h "rpf_build_chunked"
  pure (GpuForest fh xs rvss ntrees maxd)

-- | 'Data.RPTree.Batch.forestBatch' (Batch.hs:48-63).
forestBatch :: Word64 -> Int -> Int -> Int -> Double -> Int -> V.Vector (Embed DVector Double x) -> GpuForest x
forestBatch seed maxd minl ntrees pnz dim xs = unsafePerformIO (buildWith [0] Nothing seed maxd minl ntrees pnz dim xs)
{-# NOINLINE forestBatch #-}

-- | 'forestBatch' on an explicit list of GPUs (e.g. @[0 .. 7]@: 8 B200 of one box, the trees sharded across them).
forestOn :: [Int] -> Word64 -> Int -> Int -> Int -> Double -> Int -> V.Vector (Embed DVector Double x) -> IO (GpuForest x)
forestOn devs = buildWith devs Nothing

-- | 'Data.RPTree.Batch.treeBatch' (Batch.hs:29-41).
treeBatch :: Word64 -> Int -> Int -> Double -> Int -> V.Vector (Embed DVector Double x) -> GpuForest x
treeBatch seed maxd minl = forestBatch seed maxd minl 1

-- | 'Data.RPTree.Conduit.forest' (Conduit.hs:104-121), same argument order: the source is drained into a vector and the
-- engine replays the conduit's @chunksOf chunksize@ fold itself (rpf_build_chunked), reproducing insert's Bin and Tip
-- cases per chunk (Internal.hs:243-297).
forest :: Monad m => Word64 -> Int -> Int -> Int -> Int -> Double -> Int -> C.ConduitT () (Embed DVector Double x) m () -> m (GpuForest x)
forest seed maxd minl ntrees chunksize pnz dim src = do
  xs <- C.runConduit (src C..| CC.sinkVector)
  pure $! unsafePerformIO (buildWith [0] (Just chunksize) seed maxd minl ntrees pnz dim xs)

-- | 'forest' without draining the source first: every @chunksize@ rows the conduit yields are handed to the engine as they
-- arrive ('c_insChunk'), exactly the reference's @chunksOf n .| foldl insertMulti@ (Conduit.hs:157-176).  The handle holds a
-- complete forest after every chunk, so a consumer may query between chunks; memory on the device grows geometrically.
forestStreaming :: MonadIO m => Word64 -> Int -> Int -> Int -> Int -> Double -> Int -> C.ConduitT () (Embed DVector Double x) m () -> m (GpuForest x)
forestStreaming seed maxd minl ntrees chunksize pnz dim src = do
  let rvss = sample seed $ do                                   -- Conduit.hs:116-118, verbatim
        rvs <- replicateM ntrees $ V.replicateM maxd (sparse pnz dim stdNormal)
        pure $ IM.fromList $ zip [0 ..] rvs
      (off, idx, val) = csrOf rvss
  fh <- liftIO $ withDevices [0] pure
  liftIO $ withForeignPtr fh $ \h -> do
    withArray off $ \po -> withArray idx $ \pi' -> withArray val $ \pv ->
      c_setHp h (fromIntegral ntrees) (fromIntegral maxd) po pi' pv >>= check h "rpf_set_hyperplanes"
    c_insBegin h (fromIntegral dim) (fromIntegral maxd) (fromIntegral minl) >>= check h "rpf_insert_begin"
  let insert1 chunk = liftIO $ withForeignPtr fh $ \h -> do
        let xs = V.fromList chunk
        VS.unsafeWith (packRows dim xs) $ \p -> c_insChunk h p (fromIntegral (V.length xs)) >>= check h "rpf_insert_chunk"
        pure xs
  xss <- C.runConduit (src C..| CL.chunksOf chunksize C..| CC.mapM insert1 C..| CC.sinkList)
  liftIO $ withForeignPtr fh $ \h -> c_insEnd h >>= check h "rpf_insert_end"
  pure (GpuForest fh (V.concat xss) rvss ntrees maxd)

queryPtr :: DVector Double -> (Ptr CDouble -> IO a) -> IO a
queryPtr (DV q) = VS.unsafeWith (VS.map realToFrac (VS.convert q))

-- | 'Data.RPTree.knn' (RPTree.hs:168-176), same argument order.  The engine implements @distf = metricL2@ only; the
-- argument is used to re-evaluate the distance of the first result on the Haskell side and must agree with the engine's.
knn :: (DVector Double -> DVector Double -> Double) -> Int -> GpuForest x -> DVector Double -> V.Vector (Double, Embed DVector Double x)
knn = knnWith 0
-- | 'Data.RPTree.knnPQ' (RPTree.hs:181-194), same argument order; @distf@ as in 'knn'.
knnPQ :: (DVector Double -> DVector Double -> Double) -> Int -> GpuForest x -> DVector Double -> V.Vector (Double, Embed DVector Double x)
knnPQ = knnWith 1

knnWith :: Int32 -> (DVector Double -> DVector Double -> Double) -> Int -> GpuForest x -> DVector Double -> V.Vector (Double, Embed DVector Double x)
knnWith dedup distf k gf q = unsafePerformIO $ withForeignPtr (gfHandle gf) $ \h -> queryPtr q $ \pq -> alloca $ \pc -> do
  pdv <- mallocForeignPtrArray k
  piv <- mallocForeignPtrArray k
  withForeignPtr pdv $ \pd -> withForeignPtr piv $ \pi' ->
    c_knn h pq 1 (fromIntegral k) dedup pd pi' pc >>= check h "rpf_knn"
  m <- fromIntegral <$> peek pc
  let ds = VS.unsafeFromForeignPtr0 pdv m :: VS.Vector CDouble
      is = VS.unsafeFromForeignPtr0 piv m :: VS.Vector Word32
      res = V.generate m (\j -> (realToFrac (ds VS.! j), gfRows gf V.! fromIntegral (is VS.! j)))
  when (m > 0) $ do
    let (d0, x0) = V.head res
        dref = distf (eEmbed x0) q
    when (abs (dref - d0) > 4 * abs d0 * 2.220446049250313e-16) $
      throwIO (ErrorCall "Data.RPTree.CUDA.knn: the engine implements metricL2 only")
  pure res

-- | 'Data.RPTree.knnH' with distf = metricL2 (RPTree.hs:199-217): whole leaves in margin-priority order, prepended
-- while the running total stays <= k.  Not sorted by distance, not cut to k -- as in the reference.
knnH :: Int -> GpuForest x -> DVector Double -> V.Vector (Double, Embed DVector Double x)
knnH k gf q = unsafePerformIO $ withForeignPtr (gfHandle gf) $ \h -> queryPtr q $ \pq -> do
  cap <- fromIntegral <$> c_knnHCap h (fromIntegral k)
  allocaArray cap $ \pd -> allocaArray cap $ \pi' -> alloca $ \pc -> do
    c_knnH h pq nullPtr 1 (fromIntegral k) (fromIntegral cap) pd pi' pc >>= check h "rpf_knn_h"
    m <- fromIntegral <$> peek pc
    ds <- peekArray m pd
    is <- peekArray m pi'
    pure $ V.fromList [ (realToFrac d, gfRows gf V.! fromIntegral i) | (d, i) <- zip ds is ]

-- | Engine-side checkpoint (the CBOR form of 'serialiseRPForest' is still available through 'toRPForest').
saveForest :: GpuForest x -> FilePath -> IO ()
saveForest gf path = withForeignPtr (gfHandle gf) $ \h -> withCString path $ \p -> c_save h p 1 >>= check h "rpf_forest_save"

-- | Data points as 'SVector's (Internal.hs:92-97): the (Int, Double) pairs are already SoA, so the rows go over as CSR.
-- Distances then follow metricSDL2 / metricSSL2 (Internal.hs:389-400) including their early-stop behaviour.
setPointsSparse :: Ptr RpfHandle -> Int -> V.Vector (SVector Double) -> IO ()
setPointsSparse h dim svs =
  withArray off $ \po -> withArray idx $ \pi' -> withArray val $ \pv ->
    c_setPtsS h (fromIntegral (V.length svs)) (fromIntegral dim) po pi' pv >>= check h "rpf_set_points_sparse"
  where
    rows = map (VU.toList . svVec) (V.toList svs)
    off  = scanl (+) 0 (map (fromIntegral . length) rows) :: [Int64]
    idx  = concatMap (map (fromIntegral . fst)) rows :: [Int32]
    val  = concatMap (map (realToFrac . snd)) rows :: [CDouble]

-- | 'Data.RPTree.candidates' for tree @t@ of the forest (RPTree.hs:293-314).
candidates :: GpuForest x -> Int -> DVector Double -> V.Vector (Embed DVector Double x)
candidates gf t q = unsafePerformIO $ withForeignPtr (gfHandle gf) $ \h -> queryPtr q $ \pq ->
  allocaArray 2 $ \poff -> do
    c_candCnt h pq 1 (fromIntegral t) poff >>= check h "rpf_candidates_count"
    [_, m] <- map fromIntegral <$> peekArray 2 poff
    allocaArray (max 1 m) $ \pids -> do
      c_cand h pq 1 (fromIntegral t) poff pids >>= check h "rpf_candidates"
      V.map ((gfRows gf V.!) . fromIntegral) . V.fromList <$> peekArray m pids

-- | 'Data.RPTree.recallWith' with distf = metricL2 (RPTree.hs:259-268).
recallWith :: GpuForest x -> Int -> DVector Double -> Double
recallWith gf k q = unsafePerformIO $ withForeignPtr (gfHandle gf) $ \h -> queryPtr q $ \pq -> alloca $ \pr -> do
  c_recall h pq 1 (fromIntegral k) pr >>= check h "rpf_recall"
  (/ fromIntegral (gfNTrees gf)) . realToFrac <$> peek pr

-- | Rebuild the pure 'RPForest' value (Bin/Tip, Internal.hs:139-148) from the flat arrays: the topology comes from
-- rpf_topology (BFS ids; child g = left child, +1 = right), thresholds/margins/leaf contents from rpf_tree_export.
toRPForest :: GpuForest x -> IO (RPForest Double (V.Vector (Embed DVector Double x)))
toRPForest gf = withForeignPtr (gfHandle gf) $ \h -> do
  nn <- fromIntegral <$> c_numNodes h
  let n = V.length (gfRows gf)
      nt = gfNTrees gf
  -- topology once (the same for every tree), then the whole forest in ONE call: [T][nodes] / [T][n] arrays
  childFp <- mallocForeignPtrArray nn; startFp <- mallocForeignPtrArray nn; sizeFp <- mallocForeignPtrArray nn
  withForeignPtr childFp $ \pc -> withForeignPtr startFp $ \ps -> withForeignPtr sizeFp $ \pz ->
    c_topology h pc nullPtr ps pz >>= check h "rpf_topology"
  let child = VS.unsafeFromForeignPtr0 childFp nn :: VS.Vector Int64
      start = VS.unsafeFromForeignPtr0 startFp nn :: VS.Vector Int64
      size  = VS.unsafeFromForeignPtr0 sizeFp nn :: VS.Vector Int64
  thrFp <- mallocForeignPtrArray (nt * nn); mloFp <- mallocForeignPtrArray (nt * nn); mhiFp <- mallocForeignPtrArray (nt * nn)
  permFp <- mallocForeignPtrArray (max 1 (nt * n))
  withForeignPtr thrFp $ \pt -> withForeignPtr mloFp $ \pl -> withForeignPtr mhiFp $ \ph -> withForeignPtr permFp $ \pp ->
    c_exportAll h pt pl ph pp >>= check h "rpf_forest_export"
  let thr  = VS.unsafeFromForeignPtr0 thrFp (nt * nn) :: VS.Vector CDouble
      mlo  = VS.unsafeFromForeignPtr0 mloFp (nt * nn) :: VS.Vector CDouble
      mhi  = VS.unsafeFromForeignPtr0 mhiFp (nt * nn) :: VS.Vector CDouble
      perm = VS.unsafeFromForeignPtr0 permFp (nt * n) :: VS.Vector Word32
      tree t = RPTree (gfVectors gf IM.! t) (go 0)
        where
          go :: Int -> RPT Double () (V.Vector (Embed DVector Double x))
          go g | c < 0     = Tip () (V.generate (fromIntegral (size VS.! g))
                                       (\i -> gfRows gf V.! fromIntegral (perm VS.! (t * n + fromIntegral (start VS.! g) + i))))
               | otherwise = Bin () (realToFrac (thr VS.! (t * nn + g)))
                                 (Margin (Max (realToFrac (mlo VS.! (t * nn + g)))) (Min (realToFrac (mhi VS.! (t * nn + g)))))
                                 (go (fromIntegral c)) (go (fromIntegral c + 1))
            where c = child VS.! g
  pure (IM.fromList [ (t, tree t) | t <- [0 .. nt - 1] ])
