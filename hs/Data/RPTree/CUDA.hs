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
  ( GpuForest, forestBatch, treeBatch, forest, knn, knnPQ, knnH, candidates, recallWith, toRPForest, withDevice
  , saveForest, setPointsSparse
  ) where

import Control.Exception (throwIO, ErrorCall(..))
import Control.Monad (replicateM, when, forM)
import Data.Int (Int32, Int64)
import Data.Word (Word32, Word64)
import Foreign.C.String (CString, peekCString, withCString)
import Foreign.C.Types (CInt(..), CDouble(..))
import Foreign.ForeignPtr (ForeignPtr, newForeignPtr, withForeignPtr)
import Foreign.Marshal.Alloc (alloca)
import Foreign.Marshal.Array (allocaArray, peekArray, withArrayLen, withArray)
import Foreign.Ptr (Ptr, FunPtr, nullPtr)
import Foreign.Storable (peek)
import System.IO.Unsafe (unsafePerformIO)

import qualified Data.IntMap.Strict as IM
import qualified Data.Vector as V
import qualified Data.Vector.Storable as VS
import qualified Data.Vector.Unboxed as VU
import System.Random.SplitMix.Distributions (sample, stdNormal)

import Data.RPTree.Gen (sparse)
import Data.RPTree.Internal (RPTree(..), RPT(..), RPForest, Embed(..), DVector(..), SVector(..), Margin(..))
import Data.Semigroup (Max(..), Min(..))

data RpfHandle

foreign import ccall safe "rpf_create"            c_create   :: Ptr (Ptr RpfHandle) -> CInt -> IO CInt
foreign import ccall safe "&rpf_destroy"          p_destroy  :: FunPtr (Ptr RpfHandle -> IO ())
foreign import ccall safe "rpf_last_error"        c_lastErr  :: Ptr RpfHandle -> IO CString
foreign import ccall safe "rpf_set_points"        c_setPts   :: Ptr RpfHandle -> Ptr CDouble -> Int64 -> Int32 -> IO CInt
foreign import ccall safe "rpf_set_hyperplanes"   c_setHp    :: Ptr RpfHandle -> Int32 -> Int32 -> Ptr Int64 -> Ptr Int32 -> Ptr CDouble -> IO CInt
foreign import ccall safe "rpf_build"             c_build    :: Ptr RpfHandle -> Int32 -> Int32 -> IO CInt
foreign import ccall safe "rpf_build_from_host"   c_buildH   :: Ptr RpfHandle -> Ptr CDouble -> Int64 -> Int32 -> Int32 -> Int32 -> IO CInt
foreign import ccall safe "rpf_build_chunked"     c_buildCh  :: Ptr RpfHandle -> Int32 -> Int32 -> Int64 -> IO CInt
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

-- | A forest living on one B200.  The payloads stay on the Haskell side, addressed by row number.
data GpuForest x = GpuForest
  { gfHandle  :: !(ForeignPtr RpfHandle)
  , gfRows    :: !(V.Vector (Embed DVector Double x))   -- ^ row id -> original item
  , gfVectors :: !(IM.IntMap (V.Vector (SVector Double))) -- ^ rvss, as drawn by 'sample'
  , gfNTrees  :: !Int
  , gfMaxD    :: !Int
  }

withDevice :: Int -> (ForeignPtr RpfHandle -> IO a) -> IO a
withDevice dev k = alloca $ \pp -> do
  rc <- c_create pp (fromIntegral dev)
  when (rc /= 0) $ throwIO (ErrorCall "rpf_create: no usable CUDA device (there is no CPU fallback)")
  h <- peek pp >>= newForeignPtr p_destroy
  k h

check :: Ptr RpfHandle -> String -> CInt -> IO ()
check h what rc = when (rc /= 0) $ do
  msg <- c_lastErr h >>= peekCString
  throwIO (ErrorCall (what ++ ": " ++ msg))

-- | One pinned, row-major n x d buffer out of the unpinned VU.Vectors (Internal.hs:122).
packRows :: V.Vector (Embed DVector Double x) -> VS.Vector CDouble
packRows = VS.concat . map (VS.map realToFrac . VS.convert . dvVec . eEmbed) . V.toList

-- | CSR over (tree-major, level-minor) of rvss; SVector's VU.Vector (Int, Double) is already SoA (Internal.hs:92-93).
csrOf :: IM.IntMap (V.Vector (SVector Double)) -> ([Int64], [Int32], [CDouble])
csrOf rvss = (scanl (+) 0 (map (fromIntegral . VU.length . svVec) svs), concatMap (map (fromIntegral . fst) . VU.toList . svVec) svs
             , concatMap (map (realToFrac . snd) . VU.toList . svVec) svs)
  where svs = concatMap V.toList (IM.elems rvss)

buildWith :: Maybe Int -> Word64 -> Int -> Int -> Int -> Double -> Int -> V.Vector (Embed DVector Double x) -> IO (GpuForest x)
buildWith chunk seed maxd minl ntrees pnz dim xs = withDevice 0 $ \fh -> withForeignPtr fh $ \h -> do
  let rvss = sample seed $ do                                   -- Batch.hs:59-61 / Conduit.hs:116-118, verbatim
        rvs <- replicateM ntrees $ V.replicateM maxd (sparse pnz dim stdNormal)
        pure $ IM.fromList $ zip [0 ..] rvs
      (off, idx, val) = csrOf rvss
      buf = packRows xs
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
forestBatch seed maxd minl ntrees pnz dim xs = unsafePerformIO (buildWith Nothing seed maxd minl ntrees pnz dim xs)
{-# NOINLINE forestBatch #-}

-- | 'Data.RPTree.Batch.treeBatch' (Batch.hs:29-41).
treeBatch :: Word64 -> Int -> Int -> Double -> Int -> V.Vector (Embed DVector Double x) -> GpuForest x
treeBatch seed maxd minl = forestBatch seed maxd minl 1

-- | 'Data.RPTree.Conduit.forest' (Conduit.hs:104-121) after the source has been drained into a vector: the engine replays
-- the conduit's @chunksOf chunksize@ fold itself (rpf_build_chunked), reproducing insert's Bin and Tip cases per chunk.
forest :: Word64 -> Int -> Int -> Int -> Int -> Double -> Int -> V.Vector (Embed DVector Double x) -> IO (GpuForest x)
forest seed maxd minl ntrees chunksize = buildWith (Just chunksize) seed maxd minl ntrees

queryPtr :: DVector Double -> (Ptr CDouble -> IO a) -> IO a
queryPtr (DV q) = VS.unsafeWith (VS.map realToFrac (VS.convert q))

-- | 'Data.RPTree.knn' with distf = metricL2 (RPTree.hs:168-176).
knn :: Int -> GpuForest x -> DVector Double -> V.Vector (Double, Embed DVector Double x)
knn = knnWith 0
-- | 'Data.RPTree.knnPQ' with distf = metricL2 (RPTree.hs:181-194).
knnPQ :: Int -> GpuForest x -> DVector Double -> V.Vector (Double, Embed DVector Double x)
knnPQ = knnWith 1

knnWith :: Int32 -> Int -> GpuForest x -> DVector Double -> V.Vector (Double, Embed DVector Double x)
knnWith dedup k gf q = unsafePerformIO $ withForeignPtr (gfHandle gf) $ \h -> queryPtr q $ \pq ->
  allocaArray k $ \pd -> allocaArray k $ \pi' -> alloca $ \pc -> do
    c_knn h pq 1 (fromIntegral k) dedup pd pi' pc >>= check h "rpf_knn"
    m <- fromIntegral <$> peek pc
    ds <- peekArray m pd
    is <- peekArray m pi'
    pure $ V.fromList [ (realToFrac d, gfRows gf V.! fromIntegral i) | (d, i) <- zip ds is ]

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
  (child, start, size) <- allocaArray nn $ \pc -> allocaArray nn $ \ps -> allocaArray nn $ \pz -> do
    c_topology h pc nullPtr ps pz >>= check h "rpf_topology"
    (,,) <$> (VU.fromList <$> peekArray nn pc) <*> (VU.fromList <$> peekArray nn ps) <*> (VU.fromList <$> peekArray nn pz)
  trees <- forM [0 .. gfNTrees gf - 1] $ \t ->
    allocaArray nn $ \pt -> allocaArray nn $ \pl -> allocaArray nn $ \ph -> allocaArray (max 1 n) $ \pp -> do
      c_export h (fromIntegral t) pt pl ph pp >>= check h "rpf_tree_export"
      thr <- VU.fromList . map realToFrac <$> peekArray nn pt
      mlo <- VU.fromList . map realToFrac <$> peekArray nn pl
      mhi <- VU.fromList . map realToFrac <$> peekArray nn ph
      perm <- VU.fromList <$> peekArray n pp
      let go :: Int -> RPT Double () (V.Vector (Embed DVector Double x))
          go g | c < 0     = Tip () (V.generate (fromIntegral (size VU.! g)) (\i -> gfRows gf V.! fromIntegral (perm VU.! (fromIntegral (start VU.! g) + i))))
               | otherwise = Bin () (thr VU.! g) (Margin (Max (mlo VU.! g)) (Min (mhi VU.! g))) (go (fromIntegral c)) (go (fromIntegral c + 1))
            where c = (child VU.! g) :: Int64
      pure (t, RPTree (gfVectors gf IM.! t) (go 0))
  pure (IM.fromList trees)
