(ns org.soulspace.qclojure.domain.math.b200.backend
  "P2: complex linear-algebra backend over libqcb200.so (`qcb_la_*`, include/qcb200.h), implementing the five protocols of
  domain/math/protocols.clj (BackendAdapter :12-78, MatrixAlgebra :81-355, MatrixDecompositions :357-444, MatrixFunctions
  :446-492, MatrixAnalysis :494-521) the way fastmath/backend.clj:84-245 does for the default backend.

  NOT COMPILED OR RUN IN THIS REPOSITORY'S CI (no JVM in the build image, DESIGN.md §1).  The same entry points are driven
  from Python by qclojure_b200/linalg.py (`B200ComplexBackend`, method names = protocol names in snake_case) and checked
  against NumPy / SciPy there (tests/test_la_host.py, tests/test_gpu_parity.py::test_p2_linear_algebra_ops).

  Backend representation: row-major interleaved doubles, {:rows r :cols c :data double[2rc]} for matrices and
  {:n n :data double[2n]} for vectors; scalars stay fastmath Vec2.  Install with
    (cla/with-backend (b200/create-backend) ...)            ; complex_linear_algebra.clj:131-139 accepts an instance
  (the keyword registry at :40 is private; `set-backend! :b200` needs the one-line upstream patch INTEGRATION.md §5 names)."
  (:require [org.soulspace.qclojure.domain.math.protocols :as proto]
            [fastmath.complex :as fc])
  (:import (java.lang.foreign Arena FunctionDescriptor Linker MemoryLayout MemorySegment SymbolLookup ValueLayout)
           (java.lang.invoke MethodHandle)))

;;; ------------------------------------------------------------------ FFM plumbing (same as adapter/backend/b200_simulator.clj)
(def ^:private ^Linker linker (Linker/nativeLinker))
(defonce ^:private lookup
  (delay (SymbolLookup/libraryLookup (or (System/getProperty "qcb200.lib") "libqcb200.so") (Arena/global))))
(def ^:private I32 ValueLayout/JAVA_INT)
(def ^:private I64 ValueLayout/JAVA_LONG)
(def ^:private F64 ValueLayout/JAVA_DOUBLE)
(def ^:private PTR ValueLayout/ADDRESS)

(def ^:private handle-cache (atom {}))
(defn- ^MethodHandle ffi [name & arg-layouts]
  (or (@handle-cache name)
      (let [h (.downcallHandle linker (.orElseThrow (.find ^SymbolLookup @lookup name))
                               (FunctionDescriptor/of I32 (into-array MemoryLayout arg-layouts)))]
        (swap! handle-cache assoc name h)
        h)))

(defn- call!
  "Invoke `int32_t name(qcb_handle h, args...)`; h = the device handle of the backend (NULL = host-only entry points)."
  [name layouts & args]
  (let [rc (.invokeWithArguments (apply ffi name layouts) ^java.util.List (vec args))]
    (when-not (zero? (int rc))
      (throw (ex-info (str name " failed") {:qcb-status rc})))))

(defn- seg-of ^MemorySegment [^Arena a ^doubles d] (.allocateFrom a F64 d))
(defn- out-seg ^MemorySegment [^Arena a n-doubles] (.allocate a (long (* 8 (max 1 n-doubles))) 8))
(defn- doubles-of ^doubles [^MemorySegment s n-doubles] (.toArray (.asSlice s 0 (long (* 8 n-doubles))) F64))

;;; ------------------------------------------------------------------ representation
(defn- mat [r c ^doubles d] {:rows r :cols c :data d})
(defn- vec* [n ^doubles d] {:n n :data d})
(defn- vec2s->doubles ^doubles [zs]
  (let [d (double-array (* 2 (count zs)))]
    (doseq [[i z] (map-indexed vector zs)]
      (aset d (* 2 i) (double (fc/re z))) (aset d (inc (* 2 i)) (double (fc/im z))))
    d))
(defn- doubles->vec2s [^doubles d]
  (mapv (fn [i] (fc/complex (aget d (* 2 i)) (aget d (inc (* 2 i))))) (range (quot (alength d) 2))))
(defn- ensure-complex [s] (if (number? s) (fc/complex (double s) 0.0) s))

(defrecord B200ComplexBackend [tolerance config handle])   ; handle: MemorySegment of a qcb_handle, or MemorySegment/NULL

(defn create-backend
  "opts: {:tolerance 1e-12 :handle <qcb_handle MemorySegment>}.  Without a handle only the host-side entry points
  (decompositions, matrix functions, predicates, transpose, hadamard, solve, inverse) are usable; the products
  (matvec, matmul, kron, inner, outer, trace, norm2, axpby) run on the GPU and need the handle of a simulator
  (b200_simulator/open-handle)."
  ([] (create-backend {}))
  ([opts] (->B200ComplexBackend (or (:tolerance opts) 1.0e-12) opts (or (:handle opts) MemorySegment/NULL))))

;;; ------------------------------------------------------------------ BackendAdapter (protocols.clj:12-78)
(extend-protocol proto/BackendAdapter
  B200ComplexBackend
  (vector->backend [_ v]
    (cond (and (map? v) (:data v)) v
          (vector? v) (vec* (count v) (vec2s->doubles (map ensure-complex v)))
          :else (vec* 1 (vec2s->doubles [(ensure-complex v)]))))
  (backend->vector [_ v] (if (map? v) (doubles->vec2s (:data v)) v))
  (matrix->backend [_ m]
    (cond (and (map? m) (:data m)) m
          (and (map? m) (:real m)) (mat (count (:real m)) (count (first (:real m)))
                                        (vec2s->doubles (mapcat (fn [rr ir] (map fc/complex rr ir)) (:real m) (:imag m))))
          :else (mat (count m) (count (first m)) (vec2s->doubles (map ensure-complex (apply concat m))))))
  (backend->matrix [_ m]
    (if (map? m)
      (let [zs (doubles->vec2s (:data m))] (mapv vec (partition (:cols m) zs)))
      m))
  (scalar->backend [_ s] (ensure-complex s))
  (backend->scalar [_ s] s))

;;; ------------------------------------------------------------------ helpers over the C entry points
(defn- binary-same-shape
  "out = f(A, B) elementwise through qcb_la_axpby(alpha, A, beta, B)."
  [b A B [ar ai] [br bi]]
  (with-open [a (Arena/ofConfined)]
    (let [n (* (:rows A 1) (:cols A (:n A)))
          out (out-seg a (* 2 n))]
      (call! "qcb_la_axpby" [PTR PTR PTR PTR PTR I64 PTR] (:handle b)
             (seg-of a (double-array [ar ai])) (seg-of a (:data A)) (seg-of a (double-array [br bi])) (seg-of a (:data B))
             (long n) out)
      (assoc A :data (doubles-of out (* 2 n))))))

(defn- square-out
  "n x n -> n x n host-side function (inverse, matrix functions)."
  [b name A]
  (with-open [a (Arena/ofConfined)]
    (let [n (:rows A) out (out-seg a (* 2 n n))]
      (call! name [PTR PTR I64 PTR] (:handle b) (seg-of a (:data A)) (long n) out)
      (mat n n (doubles-of out (* 2 n n))))))

(defn- predicate [b name A eps]
  (with-open [a (Arena/ofConfined)]
    (let [flag (.allocate a 4 4)]
      (call! name [PTR PTR I64 F64 PTR] (:handle b) (seg-of a (:data A)) (long (:rows A)) (double eps) flag)
      (pos? (.get flag I32 0)))))

(defn- complex-out [b name layouts & args]
  (with-open [a (Arena/ofConfined)]
    (let [out (out-seg a 2)]
      (apply call! name layouts (:handle b) (concat (map #(if (fn? %) (% a) %) args) [out]))
      (fc/complex (.getAtIndex out F64 0) (.getAtIndex out F64 1)))))

;;; ------------------------------------------------------------------ MatrixAlgebra (protocols.clj:81-355)
(extend-protocol proto/MatrixAlgebra
  B200ComplexBackend
  (shape [_ A] [(:rows A) (:cols A)])
  (add [b A B] (binary-same-shape b A B [1.0 0.0] [1.0 0.0]))
  (subtract [b A B] (binary-same-shape b A B [1.0 0.0] [-1.0 0.0]))
  (scale [b A alpha] (let [z (ensure-complex alpha)] (binary-same-shape b A A [(fc/re z) (fc/im z)] [0.0 0.0])))
  (negate [b A] (binary-same-shape b A A [-1.0 0.0] [0.0 0.0]))
  (matrix-multiply [b A B]
    (with-open [a (Arena/ofConfined)]
      (let [m (:rows A) k (:cols A) n (:cols B) out (out-seg a (* 2 m n))]
        (call! "qcb_la_matmul" [PTR PTR PTR I64 I64 I64 PTR] (:handle b) (seg-of a (:data A)) (seg-of a (:data B))
               (long m) (long k) (long n) out)
        (mat m n (doubles-of out (* 2 m n))))))
  (matrix-vector-product [b A x]
    (with-open [a (Arena/ofConfined)]
      (let [r (:rows A) out (out-seg a (* 2 r))]
        (call! "qcb_la_matvec" [PTR PTR PTR I64 I64 PTR] (:handle b) (seg-of a (:data A)) (seg-of a (:data x))
               (long r) (long (:cols A)) out)
        (vec* r (doubles-of out (* 2 r))))))
  (inner-product [b x y]                                     ; conjugates its FIRST argument (fastmath backend :185-197)
    (complex-out b "qcb_la_inner" [PTR PTR PTR I64 PTR] #(seg-of % (:data x)) #(seg-of % (:data y)) (long (:n x))))
  (outer-product [b x y]                                     ; x y^H
    (with-open [a (Arena/ofConfined)]
      (let [n (:n x) m (:n y) out (out-seg a (* 2 n m))]
        (call! "qcb_la_outer" [PTR PTR PTR I64 I64 PTR] (:handle b) (seg-of a (:data x)) (seg-of a (:data y)) (long n) (long m) out)
        (mat n m (doubles-of out (* 2 n m))))))
  (hadamard-product [b A B]
    (with-open [a (Arena/ofConfined)]
      (let [n (* (:rows A) (:cols A)) out (out-seg a (* 2 n))]
        (call! "qcb_la_hadamard" [PTR PTR PTR I64 PTR] (:handle b) (seg-of a (:data A)) (seg-of a (:data B)) (long n) out)
        (assoc A :data (doubles-of out (* 2 n))))))
  (kronecker-product [b A B]
    (with-open [a (Arena/ofConfined)]
      (let [r (* (:rows A) (:rows B)) c (* (:cols A) (:cols B)) out (out-seg a (* 2 r c))]
        (call! "qcb_la_kron" [PTR PTR I64 I64 PTR I64 I64 PTR] (:handle b) (seg-of a (:data A)) (long (:rows A)) (long (:cols A))
               (seg-of a (:data B)) (long (:rows B)) (long (:cols B)) out)
        (mat r c (doubles-of out (* 2 r c))))))
  (transpose [b A]
    (with-open [a (Arena/ofConfined)]
      (let [out (out-seg a (* 2 (:rows A) (:cols A)))]
        (call! "qcb_la_transpose" [PTR PTR I64 I64 I32 PTR] (:handle b) (seg-of a (:data A)) (long (:rows A)) (long (:cols A)) (int 0) out)
        (mat (:cols A) (:rows A) (doubles-of out (* 2 (:rows A) (:cols A)))))))
  (conjugate-transpose [b A]
    (with-open [a (Arena/ofConfined)]
      (let [out (out-seg a (* 2 (:rows A) (:cols A)))]
        (call! "qcb_la_transpose" [PTR PTR I64 I64 I32 PTR] (:handle b) (seg-of a (:data A)) (long (:rows A)) (long (:cols A)) (int 1) out)
        (mat (:cols A) (:rows A) (doubles-of out (* 2 (:rows A) (:cols A)))))))
  (trace [b A] (complex-out b "qcb_la_trace" [PTR PTR I64 PTR] #(seg-of % (:data A)) (long (:rows A))))
  (norm2 [b x]
    (with-open [a (Arena/ofConfined)]
      (let [out (out-seg a 1)]
        (call! "qcb_la_norm2" [PTR PTR I64 PTR] (:handle b) (seg-of a (:data x)) (long (:n x)) out)
        (.getAtIndex out F64 0))))
  (solve-linear-system [b A rhs]
    (with-open [a (Arena/ofConfined)]
      (let [n (:rows A) out (out-seg a (* 2 n))]
        (call! "qcb_la_solve" [PTR PTR PTR I64 I64 PTR] (:handle b) (seg-of a (:data A)) (seg-of a (:data rhs)) (long n) (long 1) out)
        (vec* n (doubles-of out (* 2 n))))))
  (inverse [b A] (square-out b "qcb_la_inverse" A))
  (hermitian? ([b A] (predicate b "qcb_la_is_hermitian" A (:tolerance b))) ([b A eps] (predicate b "qcb_la_is_hermitian" A eps)))
  (diagonal? ([b A] (predicate b "qcb_la_is_diagonal" A (:tolerance b))) ([b A eps] (predicate b "qcb_la_is_diagonal" A eps)))
  (unitary? ([b U] (predicate b "qcb_la_is_unitary" U (:tolerance b))) ([b U eps] (predicate b "qcb_la_is_unitary" U eps)))
  (positive-semidefinite? [b A] (predicate b "qcb_la_is_positive_semidefinite" A (:tolerance b))))

;;; ------------------------------------------------------------------ MatrixDecompositions (protocols.clj:357-444)
(extend-protocol proto/MatrixDecompositions
  B200ComplexBackend
  (eigen-hermitian [b A]                                     ; eigenvalues ascending, eigenvectors as rows of the output
    (with-open [a (Arena/ofConfined)]
      (let [n (:rows A) w (out-seg a n) v (out-seg a (* 2 n n))]
        (call! "qcb_la_eigen_hermitian" [PTR PTR I64 PTR PTR] (:handle b) (seg-of a (:data A)) (long n) w v)
        (let [vals (doubles-of w n) vecs (doubles-of v (* 2 n n))]
          {:eigenvalues (vec* n (double-array (mapcat (fn [x] [x 0.0]) vals)))
           :eigenvectors (mapv (fn [k] (vec* n (java.util.Arrays/copyOfRange vecs (int (* 2 n k)) (int (* 2 n (inc k)))))) (range n))}))))
  (eigen-general [b A]
    (with-open [a (Arena/ofConfined)]
      (let [n (:rows A) w (out-seg a (* 2 n)) v (out-seg a (* 2 n n))]
        (call! "qcb_la_eigen_general" [PTR PTR I64 PTR PTR] (:handle b) (seg-of a (:data A)) (long n) w v)
        (let [vecs (doubles-of v (* 2 n n))]
          {:eigenvalues (vec* n (doubles-of w (* 2 n)))
           :eigenvectors (mapv (fn [k] (vec* n (java.util.Arrays/copyOfRange vecs (int (* 2 n k)) (int (* 2 n (inc k)))))) (range n))}))))
  (svd [b A]
    (with-open [a (Arena/ofConfined)]
      (let [m (:rows A) n (:cols A) k (min m n) U (out-seg a (* 2 m m)) S (out-seg a k) Vh (out-seg a (* 2 n n))]
        (call! "qcb_la_svd" [PTR PTR I64 I64 PTR PTR PTR] (:handle b) (seg-of a (:data A)) (long m) (long n) U S Vh)
        {:U (mat m m (doubles-of U (* 2 m m)))
         :S (vec* k (double-array (mapcat (fn [x] [x 0.0]) (doubles-of S k))))
         :V† (mat n n (doubles-of Vh (* 2 n n)))})))
  (lu-decomposition [b A]
    (with-open [a (Arena/ofConfined)]
      (let [n (:rows A) P (out-seg a (* 2 n n)) L (out-seg a (* 2 n n)) U (out-seg a (* 2 n n))]
        (call! "qcb_la_lu" [PTR PTR I64 PTR PTR PTR] (:handle b) (seg-of a (:data A)) (long n) P L U)
        {:P (mat n n (doubles-of P (* 2 n n))) :L (mat n n (doubles-of L (* 2 n n))) :U (mat n n (doubles-of U (* 2 n n)))})))
  (qr-decomposition [b A]
    (with-open [a (Arena/ofConfined)]
      (let [m (:rows A) n (:cols A) Q (out-seg a (* 2 m m)) R (out-seg a (* 2 m n))]
        (call! "qcb_la_qr" [PTR PTR I64 I64 PTR PTR] (:handle b) (seg-of a (:data A)) (long m) (long n) Q R)
        {:Q (mat m m (doubles-of Q (* 2 m m))) :R (mat m n (doubles-of R (* 2 m n)))})))
  (cholesky-decomposition [b A]
    (with-open [a (Arena/ofConfined)]
      (let [n (:rows A) L (out-seg a (* 2 n n))]
        (call! "qcb_la_cholesky" [PTR PTR I64 PTR] (:handle b) (seg-of a (:data A)) (long n) L)
        {:L (mat n n (doubles-of L (* 2 n n)))}))))

;;; ------------------------------------------------------------------ MatrixFunctions / MatrixAnalysis (protocols.clj:446-521)
(extend-protocol proto/MatrixFunctions
  B200ComplexBackend
  (matrix-exp [b A] (square-out b "qcb_la_matrix_exp" A))
  (matrix-log [b A] (square-out b "qcb_la_matrix_log" A))
  (matrix-sqrt [b A] (square-out b "qcb_la_matrix_sqrt" A)))

(defn- real-out [b name A]
  (with-open [a (Arena/ofConfined)]
    (let [out (out-seg a 1)]
      (call! name [PTR PTR I64 I64 PTR] (:handle b) (seg-of a (:data A)) (long (:rows A)) (long (:cols A)) out)
      (.getAtIndex out F64 0))))

(extend-protocol proto/MatrixAnalysis
  B200ComplexBackend
  (spectral-norm [b A] (real-out b "qcb_la_spectral_norm" A))
  (condition-number [b A] (real-out b "qcb_la_condition_number" A)))
